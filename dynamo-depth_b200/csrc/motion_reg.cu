// Motion-mask sparsity regulariser (Trainer.py:388-399) without the reference's host syncs.
#include "dd_common.cuh"

namespace dd {

constexpr int MS_THREADS = 256;
constexpr int MS_CHUNKS = 8;

__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// partial[(b*MS_CHUNKS+chunk)*2 + {0,1}] = (sum softplus(prob[static]), count(static))
__global__ void __launch_bounds__(MS_THREADS) msparsity_fwd_kernel(const float* __restrict__ mag, const float* __restrict__ mag_sum,
                                                                   const float* __restrict__ prob, int B, int hw,
                                                                   float* __restrict__ partial) {
  __shared__ float sh[2][MS_THREADS / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float mean = __ldg(mag_sum) / (float)((size_t)B * hw);   // disp_mag.mean() (Trainer.py:397)
  float s = 0.f, n = 0.f;
  for (int i = chunk * MS_THREADS + threadIdx.x; i < hw; i += MS_CHUNKS * MS_THREADS) {
    const size_t o = (size_t)b * hw + i;
    if (__ldg(mag + o) < mean) {
      s += softplus(__ldg(prob + o));
      n += 1.f;
    }
  }
  s = warp_sum(s), n = warp_sum(n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[0][warp] = s, sh[1][warp] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tn = 0.f;
    for (int i = 0; i < MS_THREADS / 32; ++i) ts += sh[0][i], tn += sh[1][i];
    partial[(b * MS_CHUNKS + chunk) * 2 + 0] = ts;
    partial[(b * MS_CHUNKS + chunk) * 2 + 1] = tn;
  }
}

__global__ void msparsity_finalize_kernel(const float* __restrict__ partial, int B, float* __restrict__ out) {
  // single warp; B is small
  double total = 0.0, count = 0.0;
  int all_ok = 1;
  for (int b = threadIdx.x; b < B; b += 32) {
    double s = 0.0, n = 0.0;
    for (int c = 0; c < MS_CHUNKS; ++c) s += partial[(b * MS_CHUNKS + c) * 2], n += partial[(b * MS_CHUNKS + c) * 2 + 1];
    total += s, count += n;
    if (n <= 0.0) all_ok = 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    count += __shfl_xor_sync(0xffffffffu, count, o);
    all_ok &= __shfl_xor_sync(0xffffffffu, all_ok, o);
  }
  if (threadIdx.x == 0) {
    const float guard = all_ok ? 1.f : 0.f;   // torch.all(sum(static,(1,2,3)) > 0)  (Trainer.py:398)
    out[0] = all_ok ? (float)(total / count) : 0.f;
    out[1] = all_ok ? (float)(1.0 / count) : 0.f;
    out[2] = guard;
    out[3] = (float)count;
  }
}

__global__ void __launch_bounds__(MS_THREADS) msparsity_bwd_kernel(const float* __restrict__ mag, const float* __restrict__ mag_sum,
                                                                   const float* __restrict__ prob, const float* __restrict__ out,
                                                                   const float* __restrict__ grad_out, int B, int hw,
                                                                   float* __restrict__ grad_prob) {
  const size_t n = (size_t)B * hw;
  const float mean = __ldg(mag_sum) / (float)n;
  const float k = __ldg(grad_out) * __ldg(out + 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (__ldg(mag + i) < mean) {
      const float x = __ldg(prob + i);
      g = k / (1.f + expf(-x));   // d softplus = sigmoid
    }
    grad_prob[i] = g;
  }
}

}  // namespace dd

extern "C" {

size_t dd_msparsity_workspace_bytes(int B, int h, int w) {
  (void)h, (void)w;
  return (size_t)(B > 0 ? B : 0) * dd::MS_CHUNKS * 2 * sizeof(float);
}

int dd_msparsity_fwd(const float* mag, const float* mag_sum, const float* prob, int B, int h, int w, float* out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(mag && mag_sum && prob && out && B > 0 && h > 0 && w > 0, "dd_msparsity_fwd: bad arguments");
  if (!workspace || workspace_bytes < dd_msparsity_workspace_bytes(B, h, w)) {
    set_error("dd_msparsity_fwd: workspace too small");
    return DD_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  msparsity_fwd_kernel<<<dim3(MS_CHUNKS, B), MS_THREADS, 0, st>>>(mag, mag_sum, prob, B, h * w, partial); dd::count_launches(1);
  msparsity_finalize_kernel<<<1, 32, 0, st>>>(partial, B, out); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_msparsity_bwd(const float* mag, const float* mag_sum, const float* prob, const float* out, const float* grad_out,
                     int B, int h, int w, float* grad_prob, void* stream) {
  using namespace dd;
  DD_REQUIRE(mag && mag_sum && prob && out && grad_out && grad_prob && B > 0 && h > 0 && w > 0, "dd_msparsity_bwd: bad arguments");
  const size_t n = (size_t)B * h * w;
  const int blocks = (int)((n + MS_THREADS - 1) / MS_THREADS < 1184 ? (n + MS_THREADS - 1) / MS_THREADS : 1184);
  msparsity_bwd_kernel<<<blocks, MS_THREADS, 0, (cudaStream_t)stream>>>(mag, mag_sum, prob, out, grad_out, B, h * w, grad_prob); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
