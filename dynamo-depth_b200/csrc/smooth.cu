// Edge-aware smoothness (tools.py:311-326) for all (term, level) tasks of a step in one launch,
// plus the mean-normalisation of the disparity (Trainer.py:357-358) and their backward passes.
// Pure HBM streaming: each task reads inp once and img once (neighbour taps hit L1/L2).
#include <string.h>

#include "dd_common.cuh"

namespace dd {

constexpr int SM_THREADS = 256;
constexpr int SM_CHUNKS = 16;   // CTAs per (task, image)

struct SmoothArgs {
  dd_smooth_task t[DD_MAX_SMOOTH_TASKS];
  int ntasks;
  int maxB;
  float* means;      // [ntasks][maxB*maxC] per-(image,channel) spatial mean (mean_normalise tasks)
  float* partial;    // fwd: [ntasks*2][maxB*SM_CHUNKS]; bwd: [ntasks][maxB*maxC][SM_CHUNKS] (sum g_n * inp)
  const float* grad_sums;
  int maxBC;
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

// per-(task, image, channel) spatial mean: .mean(2, True).mean(3, True)
__global__ void smooth_mean_kernel(const __grid_constant__ SmoothArgs a) {
  __shared__ float sh[SM_THREADS / 32];
  const int ti = blockIdx.z;
  const dd_smooth_task& t = a.t[ti];
  if (!t.mean_normalise) return;
  const int bc = blockIdx.x;
  if (bc >= t.B * t.C) return;
  const float* p = t.inp + (size_t)bc * t.h * t.w;
  float acc = 0.f;
  for (int i = threadIdx.x; i < t.h * t.w; i += blockDim.x) acc += __ldg(p + i);
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) a.means[(size_t)ti * a.maxBC + bc] = s / (float)(t.h * t.w);
}

__device__ __forceinline__ float edge_weight(const float* __restrict__ img, size_t plane, size_t o0, size_t o1) {
  // exp(-mean_c |img0 - img1|)
  const float d0 = fabsf(__ldg(img + o0) - __ldg(img + o1));
  const float d1 = fabsf(__ldg(img + plane + o0) - __ldg(img + plane + o1));
  const float d2 = fabsf(__ldg(img + 2 * plane + o0) - __ldg(img + 2 * plane + o1));
  return expf(-((d0 + d1 + d2) / 3.f));
}

__global__ void __launch_bounds__(SM_THREADS) smooth_fwd_kernel(const __grid_constant__ SmoothArgs a) {
  __shared__ float sh[SM_THREADS / 32];
  const int ti = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const dd_smooth_task& t = a.t[ti];
  float sx = 0.f, sy = 0.f;
  if (b < t.B) {
    const int h = t.h, w = t.w;
    const size_t plane = (size_t)h * w;
    const float* img = t.img ? t.img + (size_t)b * 3 * plane : nullptr;
    for (int i = chunk * SM_THREADS + threadIdx.x; i < h * w; i += SM_CHUNKS * SM_THREADS) {
      const int r = i / w, c = i - r * w;
      const bool hx = c < w - 1, hy = r < h - 1;
      const float wx = (hx && img) ? edge_weight(img, plane, i, i + 1) : 1.f;
      const float wy = (hy && img) ? edge_weight(img, plane, i, i + w) : 1.f;
      for (int ch = 0; ch < t.C; ++ch) {
        const float* p = t.inp + ((size_t)b * t.C + ch) * plane;
        float k = 1.f;
        if (t.mean_normalise) k = a.means[(size_t)ti * a.maxBC + b * t.C + ch] + 1e-7f;
        const float v = t.mean_normalise ? __ldg(p + i) / k : __ldg(p + i);
        if (hx) {
          const float v1 = t.mean_normalise ? __ldg(p + i + 1) / k : __ldg(p + i + 1);
          sx += fabsf(v - v1) * wx;
        }
        if (hy) {
          const float v1 = t.mean_normalise ? __ldg(p + i + w) / k : __ldg(p + i + w);
          sy += fabsf(v - v1) * wy;
        }
      }
    }
  }
  const int nper = a.maxB * SM_CHUNKS;
  const float tx = block_sum(sx, sh);
  if (threadIdx.x == 0) a.partial[(size_t)(ti * 2 + 0) * nper + b * SM_CHUNKS + chunk] = tx;
  const float tyv = block_sum(sy, sh);
  if (threadIdx.x == 0) a.partial[(size_t)(ti * 2 + 1) * nper + b * SM_CHUNKS + chunk] = tyv;
}

__global__ void smooth_finalize_kernel(const float* __restrict__ partial, float* __restrict__ sums, int nper) {
  __shared__ double sh[8];
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nper; i += blockDim.x) acc += (double)partial[(size_t)k * nper + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tt += sh[i];
    sums[k] = (float)tt;
  }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// pass 1: gradient w.r.t. the (normalised) input; for mean_normalise tasks also the per-CTA partial
// of sum_q g_n(q) * inp(q), needed for the gradient through the mean.
__global__ void __launch_bounds__(SM_THREADS) smooth_bwd_kernel(const __grid_constant__ SmoothArgs a) {
  __shared__ float sh[SM_THREADS / 32];
  const int ti = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const dd_smooth_task& t = a.t[ti];
  if (t.grad_inp == nullptr || b >= t.B) return;
  const int h = t.h, w = t.w;
  const size_t plane = (size_t)h * w;
  const float* img = t.img ? t.img + (size_t)b * 3 * plane : nullptr;
  const float gx = __ldg(a.grad_sums + ti * 2 + 0), gy = __ldg(a.grad_sums + ti * 2 + 1);
  for (int ch = 0; ch < t.C; ++ch) {
    const float* p = t.inp + ((size_t)b * t.C + ch) * plane;
    float* go = t.grad_inp + ((size_t)b * t.C + ch) * plane;
    float k = 1.f;
    if (t.mean_normalise) k = a.means[(size_t)ti * a.maxBC + b * t.C + ch] + 1e-7f;
    float dot = 0.f;
    for (int i = chunk * SM_THREADS + threadIdx.x; i < h * w; i += SM_CHUNKS * SM_THREADS) {
      const int r = i / w, c = i - r * w;
      const float raw = __ldg(p + i);
      const float v = t.mean_normalise ? raw / k : raw;
      float g = 0.f;
      if (c < w - 1) {
        const float v1 = t.mean_normalise ? __ldg(p + i + 1) / k : __ldg(p + i + 1);
        g += gx * sgn(v - v1) * (img ? edge_weight(img, plane, i, i + 1) : 1.f);
      }
      if (c > 0) {
        const float v1 = t.mean_normalise ? __ldg(p + i - 1) / k : __ldg(p + i - 1);
        g -= gx * sgn(v1 - v) * (img ? edge_weight(img, plane, i - 1, i) : 1.f);
      }
      if (r < h - 1) {
        const float v1 = t.mean_normalise ? __ldg(p + i + w) / k : __ldg(p + i + w);
        g += gy * sgn(v - v1) * (img ? edge_weight(img, plane, i, i + w) : 1.f);
      }
      if (r > 0) {
        const float v1 = t.mean_normalise ? __ldg(p + i - w) / k : __ldg(p + i - w);
        g -= gy * sgn(v1 - v) * (img ? edge_weight(img, plane, i - w, i) : 1.f);
      }
      go[i] = g;
      dot += g * raw;
    }
    if (t.mean_normalise) {
      const float s = block_sum(dot, sh);
      if (threadIdx.x == 0) a.partial[((size_t)ti * a.maxBC + b * t.C + ch) * SM_CHUNKS + chunk] = s;
    }
  }
}

// pass 2 (mean_normalise tasks): d/d inp = g_n / k - (sum g_n*inp) / (k^2 * h*w),  k = mean + 1e-7
__global__ void __launch_bounds__(SM_THREADS) smooth_bwd_norm_kernel(const __grid_constant__ SmoothArgs a) {
  const int ti = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const dd_smooth_task& t = a.t[ti];
  if (!t.mean_normalise || t.grad_inp == nullptr || b >= t.B) return;
  const size_t plane = (size_t)t.h * t.w;
  for (int ch = 0; ch < t.C; ++ch) {
    const int bc = b * t.C + ch;
    const float k = a.means[(size_t)ti * a.maxBC + bc] + 1e-7f;
    float S = 0.f;
#pragma unroll
    for (int i = 0; i < SM_CHUNKS; ++i) S += a.partial[((size_t)ti * a.maxBC + bc) * SM_CHUNKS + i];
    const float corr = S / (k * k * (float)plane);
    float* go = t.grad_inp + (size_t)bc * plane;
    for (int i = chunk * SM_THREADS + threadIdx.x; i < (int)plane; i += SM_CHUNKS * SM_THREADS) go[i] = go[i] / k - corr;
  }
}

static int fill_args(SmoothArgs& args, const dd_smooth_task* tasks, int ntasks, void* workspace, size_t bytes, bool& any_norm) {
  DD_REQUIRE(tasks != nullptr && ntasks >= 1 && ntasks <= DD_MAX_SMOOTH_TASKS, "dd_smooth: ntasks=%d out of range", ntasks);
  memset(&args, 0, sizeof(args));
  args.ntasks = ntasks;
  any_norm = false;
  for (int i = 0; i < ntasks; ++i) {
    const dd_smooth_task& t = tasks[i];
    DD_REQUIRE(t.inp != nullptr && t.B > 0 && t.C > 0 && t.h > 1 && t.w > 1, "dd_smooth: bad task %d", i);
    args.t[i] = t;
    args.maxB = t.B > args.maxB ? t.B : args.maxB;
    args.maxBC = t.B * t.C > args.maxBC ? t.B * t.C : args.maxBC;
    any_norm |= t.mean_normalise != 0;
  }
  const size_t need = dd_smooth_workspace_bytes(tasks, ntasks);
  if (workspace == nullptr || bytes < need) {
    set_error("dd_smooth: workspace too small (%zu < %zu)", bytes, need);
    return DD_ERR_WORKSPACE;
  }
  args.means = reinterpret_cast<float*>(workspace);
  args.partial = args.means + (size_t)ntasks * args.maxBC;
  return DD_OK;
}

int smooth_fwd_impl(const dd_smooth_task* tasks, int ntasks, float* sums, void* workspace, size_t bytes, cudaStream_t st) {
  SmoothArgs args;
  bool any_norm;
  int rc = fill_args(args, tasks, ntasks, workspace, bytes, any_norm);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(sums != nullptr, "dd_smooth_fwd: sums is NULL");
  if (any_norm) { smooth_mean_kernel<<<dim3(args.maxBC, 1, ntasks), SM_THREADS, 0, st>>>(args); dd::count_launches(1); }
  smooth_fwd_kernel<<<dim3(SM_CHUNKS, args.maxB, ntasks), SM_THREADS, 0, st>>>(args); dd::count_launches(1);
  smooth_finalize_kernel<<<ntasks * 2, 256, 0, st>>>(args.partial, sums, args.maxB * SM_CHUNKS); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int smooth_bwd_impl(const dd_smooth_task* tasks, int ntasks, const float* grad_sums, void* workspace, size_t bytes,
                    cudaStream_t st) {
  SmoothArgs args;
  bool any_norm;
  int rc = fill_args(args, tasks, ntasks, workspace, bytes, any_norm);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(grad_sums != nullptr, "dd_smooth_bwd: grad_sums is NULL");
  args.grad_sums = grad_sums;
  if (any_norm) { smooth_mean_kernel<<<dim3(args.maxBC, 1, ntasks), SM_THREADS, 0, st>>>(args); dd::count_launches(1); }
  smooth_bwd_kernel<<<dim3(SM_CHUNKS, args.maxB, ntasks), SM_THREADS, 0, st>>>(args); dd::count_launches(1);
  if (any_norm) { smooth_bwd_norm_kernel<<<dim3(SM_CHUNKS, args.maxB, ntasks), SM_THREADS, 0, st>>>(args); dd::count_launches(1); }
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // namespace dd

extern "C" {

size_t dd_smooth_workspace_bytes(const dd_smooth_task* tasks, int ntasks) {
  if (!tasks || ntasks <= 0) return 0;
  size_t maxB = 0, maxBC = 0;
  for (int i = 0; i < ntasks; ++i) {
    maxB = (size_t)tasks[i].B > maxB ? tasks[i].B : maxB;
    const size_t bc = (size_t)tasks[i].B * tasks[i].C;
    maxBC = bc > maxBC ? bc : maxBC;
  }
  const size_t means = (size_t)ntasks * maxBC;
  const size_t fwd = (size_t)ntasks * 2 * maxB * dd::SM_CHUNKS;
  const size_t bwd = (size_t)ntasks * maxBC * dd::SM_CHUNKS;
  return (means + (fwd > bwd ? fwd : bwd)) * sizeof(float);
}

int dd_smooth_fwd(const dd_smooth_task* tasks, int ntasks, float* sums, void* workspace, size_t workspace_bytes,
                  void* stream) {
  return dd::smooth_fwd_impl(tasks, ntasks, sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dd_smooth_bwd(const dd_smooth_task* tasks, int ntasks, const float* grad_sums, void* workspace,
                  size_t workspace_bytes, void* stream) {
  return dd::smooth_bwd_impl(tasks, ntasks, grad_sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
