// Training-mode BatchNorm2d (+ exact GELU) of the Lite-Mono encoder over NCHW fp32 tensors (reference:
// networks/depth_encoder.py:113-122 BNGELU = nn.BatchNorm2d(eps=1e-5) -> nn.GELU(), :194/:208 DilatedConv.bn1).
// The stem's three 64-channel maps at 1/2 resolution are 250 MB each at bs32: cuDNN's NCHW kernels plus the separate GELU
// pass move them eight times forward (0.45 ms) and ten times backward (1.1 ms); here each direction is two streaming
// passes that read / write 16-byte vectors:
//   forward   pass 1  per-channel shifted sums  S1 = sum(x - k), S2 = sum((x - k)^2), k = first element of the channel
//                     (no cancellation in E[x^2] - mean^2), per-CTA partials, combined in double
//             pass 2  y = gelu((x - mean) * invstd * gamma + beta); the k = 0 CTA of a channel stores mean / invstd and
//                     updates the running statistics (momentum, unbiased variance) like nn.BatchNorm2d
//   backward  pass 1  g = grad_y * gelu'(z) (z re-derived from x), partial sums of g and g * xhat
//             pass 2  grad_x = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)); grad_gamma = sum(g * xhat),
//                     grad_beta = sum(g)
// HBM-bound: forward reads x twice and writes y once (the second read of a channel slab mostly hits L2), backward reads
// x and grad_y twice and writes grad_x once.  Deterministic (fixed partial order, no atomics).
#include "dd_common.cuh"

namespace dd {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_CHUNKS = 64;

struct BnArgs {
  const float* x;
  const float* gy;      // backward only
  const float* gamma;
  const float* beta;
  const float* mean_in;     // backward: saved mean / invstd
  const float* invstd_in;
  float* y;             // forward: y; backward: grad_x
  float* save_mean;
  float* save_invstd;
  float* running_mean;
  float* running_var;
  float* grad_gamma;
  float* grad_beta;
  float2* partial;      // [C][chunks]
  int B, C, HW, chunks;
  float eps, momentum;
  int gelu;
};

__device__ __forceinline__ float gelu_exact(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752f)); }
// d/dz [z * Phi(z)] = Phi(z) + z * phi(z)
__device__ __forceinline__ float gelu_grad(float z) {
  const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
  return fmaf(z, pdf, cdf);
}

// Walks the elements of channel c owned by CTA `chunk` in VEC-wide units (unit index = (b, p) flattened, stride = chunks * threads),
// four units per step so that every thread has four independent 16-byte loads in flight.
struct Unit4 {
  size_t o[4];
  bool ok[4];
};
template <int VEC, typename Fn>
__device__ __forceinline__ void for_each_unit4(int B, int C, int HW, int c, int chunk, int chunks, Fn fn) {
  const int hwv = HW / VEC;
  const long long total = (long long)B * hwv;
  const int stride = chunks * BN_THREADS;
  long long u = (long long)chunk * BN_THREADS + threadIdx.x;
  if (u >= total) return;
  int b = (int)(u / hwv), p = (int)(u - (long long)b * hwv);
  for (; u < total; u += 4ll * stride) {
    Unit4 q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q.ok[j] = u + (long long)j * stride < total;
      q.o[j] = q.ok[j] ? ((size_t)b * C + c) * HW + (size_t)p * VEC : q.o[0];
      p += stride;
      while (p >= hwv) p -= hwv, ++b;
    }
    fn(q);
  }
}
// one unit as a float4: VEC == 1 puts the element in .x and `fill` in the other lanes
template <int VEC>
__device__ __forceinline__ float4 ldv(const float* __restrict__ p, bool ok, float fill) {
  if (!ok) return make_float4(fill, fill, fill, fill);
  if (VEC == 4) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), fill, fill, fill);
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, bool ok, const float4& v) {
  if (!ok) return;
  if (VEC == 4) *reinterpret_cast<float4*>(p) = v;
  else *p = v.x;
}

__device__ __forceinline__ float2 block_sum2(float a, float b) {
  __shared__ float2 sh[BN_THREADS / 32];
  a = warp_sum(a), b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = make_float2(a, b);
  __syncthreads();
  float2 r = make_float2(0.f, 0.f);
  if (threadIdx.x < 32) {
    float2 v = threadIdx.x < BN_THREADS / 32 ? sh[threadIdx.x] : make_float2(0.f, 0.f);
    r.x = warp_sum(v.x), r.y = warp_sum(v.y);
  }
  return r;   // valid in warp 0
}

// sums of the channel's partials in double, broadcast to the CTA
__device__ __forceinline__ void combine_partials(const float2* __restrict__ partial, int chunks, double& s1, double& s2) {
  __shared__ double comb[2];
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < chunks; k += 32) {
      const float2 v = partial[k];
      a += (double)v.x, b += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
    if (threadIdx.x == 0) comb[0] = a, comb[1] = b;
  }
  __syncthreads();
  s1 = comb[0], s2 = comb[1];
}

template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const __grid_constant__ BnArgs a) {
  const int c = blockIdx.y, chunk = blockIdx.x;
  const float k0 = __ldg(a.x + (size_t)c * a.HW);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for_each_unit4<VEC>(a.B, a.C, a.HW, c, chunk, a.chunks, [&](const Unit4& q) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ldv<VEC>(a.x + q.o[j], q.ok[j], k0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d0 = v[j].x - k0, d1 = v[j].y - k0, d2 = v[j].z - k0, d3 = v[j].w - k0;
      s1[0] += d0, s1[1] += d1, s1[2] += d2, s1[3] += d3;
      s2[0] = fmaf(d0, d0, s2[0]), s2[1] = fmaf(d1, d1, s2[1]), s2[2] = fmaf(d2, d2, s2[2]), s2[3] = fmaf(d3, d3, s2[3]);
    }
  });
  const float2 r = block_sum2((s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]));
  if (threadIdx.x == 0) a.partial[(size_t)c * a.chunks + chunk] = r;
}

template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const __grid_constant__ BnArgs a) {
  const int c = blockIdx.y, chunk = blockIdx.x;
  double S1, S2;
  combine_partials(a.partial + (size_t)c * a.chunks, a.chunks, S1, S2);
  const double n = (double)a.B * a.HW;
  const double k0 = (double)__ldg(a.x + (size_t)c * a.HW);
  const double m1 = S1 / n;
  double var = S2 / n - m1 * m1;   // biased variance of the batch
  var = var < 0.0 ? 0.0 : var;
  const float mean = (float)(k0 + m1);
  const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
  if (chunk == 0 && threadIdx.x == 0) {
    a.save_mean[c] = mean, a.save_invstd[c] = invstd;
    if (a.running_mean) a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
    if (a.running_var) a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)(var * (n / (n > 1.0 ? n - 1.0 : 1.0)));
  }
  const float sc = invstd * (a.gamma ? __ldg(a.gamma + c) : 1.f), sh = a.beta ? __ldg(a.beta + c) : 0.f;
  const bool gelu = a.gelu != 0;
  for_each_unit4<VEC>(a.B, a.C, a.HW, c, chunk, a.chunks, [&](const Unit4& q) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ldv<VEC>(a.x + q.o[j], q.ok[j], 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 r = make_float4(fmaf(v[j].x - mean, sc, sh), fmaf(v[j].y - mean, sc, sh), fmaf(v[j].z - mean, sc, sh), fmaf(v[j].w - mean, sc, sh));
      if (gelu) r = make_float4(gelu_exact(r.x), gelu_exact(r.y), gelu_exact(r.z), gelu_exact(r.w));
      stv<VEC>(a.y + q.o[j], q.ok[j], r);
    }
  });
}

template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const __grid_constant__ BnArgs a) {
  const int c = blockIdx.y, chunk = blockIdx.x;
  const float mean = __ldg(a.mean_in + c), invstd = __ldg(a.invstd_in + c);
  const float gm = a.gamma ? __ldg(a.gamma + c) : 1.f, bt = a.beta ? __ldg(a.beta + c) : 0.f;
  const bool gelu = a.gelu != 0;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  auto one = [&](float xv, float gv, int j) {
    const float xh = (xv - mean) * invstd;
    const float g = gelu ? gv * gelu_grad(fmaf(xh, gm, bt)) : gv;
    s1[j] += g, s2[j] = fmaf(g, xh, s2[j]);
  };
  for_each_unit4<VEC>(a.B, a.C, a.HW, c, chunk, a.chunks, [&](const Unit4& q) {
    float4 v[2], g[2];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {   // two units at a time: four loads in flight, moderate register use around erff
#pragma unroll
      for (int j = 0; j < 2; ++j) v[j] = ldv<VEC>(a.x + q.o[h + j], q.ok[h + j], mean), g[j] = ldv<VEC>(a.gy + q.o[h + j], q.ok[h + j], 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j) one(v[j].x, g[j].x, 0), one(v[j].y, g[j].y, 1), one(v[j].z, g[j].z, 2), one(v[j].w, g[j].w, 3);
    }
  });
  const float2 r = block_sum2((s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]));
  if (threadIdx.x == 0) a.partial[(size_t)c * a.chunks + chunk] = r;
}

template <int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const __grid_constant__ BnArgs a) {
  const int c = blockIdx.y, chunk = blockIdx.x;
  double S1, S2;
  combine_partials(a.partial + (size_t)c * a.chunks, a.chunks, S1, S2);
  const double n = (double)a.B * a.HW;
  if (chunk == 0 && threadIdx.x == 0) {
    if (a.grad_beta) a.grad_beta[c] = (float)S1;
    if (a.grad_gamma) a.grad_gamma[c] = (float)S2;
  }
  if (a.y == nullptr) return;
  const float mg = (float)(S1 / n), mgx = (float)(S2 / n);
  const float mean = __ldg(a.mean_in + c), invstd = __ldg(a.invstd_in + c);
  const float gm = a.gamma ? __ldg(a.gamma + c) : 1.f, bt = a.beta ? __ldg(a.beta + c) : 0.f;
  const float sc = gm * invstd;
  const bool gelu = a.gelu != 0;
  auto one = [&](float xv, float gv) {
    const float xh = (xv - mean) * invstd;
    const float g = gelu ? gv * gelu_grad(fmaf(xh, gm, bt)) : gv;
    return sc * ((g - mg) - xh * mgx);
  };
  for_each_unit4<VEC>(a.B, a.C, a.HW, c, chunk, a.chunks, [&](const Unit4& q) {
    float4 v[2], g[2];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
#pragma unroll
      for (int j = 0; j < 2; ++j) v[j] = ldv<VEC>(a.x + q.o[h + j], q.ok[h + j], mean), g[j] = ldv<VEC>(a.gy + q.o[h + j], q.ok[h + j], 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        stv<VEC>(a.y + q.o[h + j], q.ok[h + j], make_float4(one(v[j].x, g[j].x), one(v[j].y, g[j].y), one(v[j].z, g[j].z), one(v[j].w, g[j].w)));
    }
  });
}

// CTAs per channel: ~8 resident CTAs per SM over all channels, at least 4 units per thread
static int bn_chunks(int B, int C, int HW, int vec) {
  const long long units = (long long)B * (HW / vec);
  long long k = (148 * 8 + C - 1) / C;
  const long long cap = (units + BN_THREADS * 4 - 1) / (BN_THREADS * 4);
  k = k > cap ? cap : k;
  k = k > BN_MAX_CHUNKS ? BN_MAX_CHUNKS : k;
  return (int)(k < 1 ? 1 : k);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int bn_check(const char* what, int B, int C, int HW) {
  DD_REQUIRE(B > 0 && C > 0 && HW > 0, "%s: bad shape B=%d C=%d HW=%d", what, B, C, HW);
  DD_REQUIRE(C <= 65535, "%s: too many channels for one launch", what);
  return DD_OK;
}

}  // namespace dd

extern "C" {

size_t dd_bn_workspace_bytes(int C) { return C > 0 ? (size_t)C * dd::BN_MAX_CHUNKS * sizeof(float2) : 0; }

int dd_bn_gelu_fwd(const float* x, int B, int C, int HW, const float* gamma, const float* beta, float eps, float momentum, int gelu,
                   float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* workspace,
                   size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && y && save_mean && save_invstd, "dd_bn_gelu_fwd: NULL pointer");
  if (int rc = bn_check("dd_bn_gelu_fwd", B, C, HW)) return rc;
  if (!workspace || workspace_bytes < dd_bn_workspace_bytes(C)) {
    set_error("dd_bn_gelu_fwd: workspace too small (%zu < %zu)", workspace_bytes, dd_bn_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  const bool v4 = HW % 4 == 0 && aligned16(x) && aligned16(y);
  BnArgs a = {};
  a.x = x, a.gamma = gamma, a.beta = beta, a.y = y, a.save_mean = save_mean, a.save_invstd = save_invstd;
  a.running_mean = running_mean, a.running_var = running_var, a.partial = reinterpret_cast<float2*>(workspace);
  a.B = B, a.C = C, a.HW = HW, a.chunks = bn_chunks(B, C, HW, v4 ? 4 : 1), a.eps = eps, a.momentum = momentum, a.gelu = gelu;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(a.chunks, C);
  if (v4) bn_stats_kernel<4><<<grid, BN_THREADS, 0, st>>>(a), bn_apply_kernel<4><<<grid, BN_THREADS, 0, st>>>(a);
  else bn_stats_kernel<1><<<grid, BN_THREADS, 0, st>>>(a), bn_apply_kernel<1><<<grid, BN_THREADS, 0, st>>>(a);
  count_launches(2);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_bn_gelu_bwd(const float* x, const float* grad_y, int B, int C, int HW, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int gelu, float* grad_x, float* grad_gamma, float* grad_beta,
                   void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && grad_y && save_mean && save_invstd, "dd_bn_gelu_bwd: NULL pointer");
  DD_REQUIRE(grad_x || grad_gamma || grad_beta, "dd_bn_gelu_bwd: no gradient requested");
  if (int rc = bn_check("dd_bn_gelu_bwd", B, C, HW)) return rc;
  if (!workspace || workspace_bytes < dd_bn_workspace_bytes(C)) {
    set_error("dd_bn_gelu_bwd: workspace too small (%zu < %zu)", workspace_bytes, dd_bn_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  const bool v4 = HW % 4 == 0 && aligned16(x) && aligned16(grad_y) && (!grad_x || aligned16(grad_x));
  BnArgs a = {};
  a.x = x, a.gy = grad_y, a.gamma = gamma, a.beta = beta, a.mean_in = save_mean, a.invstd_in = save_invstd;
  a.y = grad_x, a.grad_gamma = grad_gamma, a.grad_beta = grad_beta, a.partial = reinterpret_cast<float2*>(workspace);
  a.B = B, a.C = C, a.HW = HW, a.chunks = bn_chunks(B, C, HW, v4 ? 4 : 1), a.gelu = gelu;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(a.chunks, C);
  if (v4) bn_bwd_stats_kernel<4><<<grid, BN_THREADS, 0, st>>>(a), bn_bwd_apply_kernel<4><<<grid, BN_THREADS, 0, st>>>(a);
  else bn_bwd_stats_kernel<1><<<grid, BN_THREADS, 0, st>>>(a), bn_bwd_apply_kernel<1><<<grid, BN_THREADS, 0, st>>>(a);
  count_launches(2);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
