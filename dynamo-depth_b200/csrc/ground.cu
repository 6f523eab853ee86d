// RANSAC ground-plane hypothesis scoring (tools.py:113-139): every CTA streams a slice of one image's ground
// points and scores all hypotheses paired with that image, 20 at a time in registers.
#include "dd_common.cuh"

namespace dd {

constexpr int GS_THREADS = 256;
constexpr int GS_HYP = 20;      // hypotheses scored per pass over the points
constexpr int GS_MAX_HYP = 512; // hypotheses per image staged in shared memory

__global__ void __launch_bounds__(GS_THREADS) ground_score_kernel(const float* __restrict__ pts, const float* __restrict__ w, int B,
                                                                 int H, int W, int row0, int K, float tol,
                                                                 int* __restrict__ counts) {
  __shared__ float ws[GS_MAX_HYP * 3];
  const int img = blockIdx.y;
  const int n_hyp = (K - img + B - 1) / B;   // hypotheses k = img + B*j < K
  for (int i = threadIdx.x; i < n_hyp * 3; i += GS_THREADS) ws[i] = __ldg(w + (size_t)(img + B * (i / 3)) * 3 + (i % 3));
  __syncthreads();
  const size_t P = (size_t)H * W;
  const float* px = pts + (size_t)img * 3 * P;
  const int n0 = row0 * W, n1 = H * W;
  for (int j0 = 0; j0 < n_hyp; j0 += GS_HYP) {
    int cnt[GS_HYP];
#pragma unroll
    for (int j = 0; j < GS_HYP; ++j) cnt[j] = 0;
    for (int n = n0 + blockIdx.x * GS_THREADS + threadIdx.x; n < n1; n += gridDim.x * GS_THREADS) {
      const float x = __ldg(px + n), y = __ldg(px + P + n), z = __ldg(px + 2 * P + n);
#pragma unroll
      for (int j = 0; j < GS_HYP; ++j) {
        if (j0 + j < n_hyp) {
          const float* wj = ws + (j0 + j) * 3;
          const float dlt = (x * wj[0] + z * wj[1] + wj[2]) - y;   // [x, z, 1] @ w - y   (tools.py:101-110,155-164)
          cnt[j] += fabsf(dlt) < tol ? 1 : 0;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < GS_HYP; ++j) {
      int v = cnt[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && j0 + j < n_hyp && v) atomicAdd(counts + img + B * (j0 + j), v);
    }
  }
}

}  // namespace dd

extern "C" int dd_ground_score(const float* points, const float* w, int B, int H, int W, int row0, int K, float tol,
                               int32_t* counts, void* stream) {
  using namespace dd;
  DD_REQUIRE(points && w && counts && B > 0 && H > 0 && W > 0 && K > 0 && row0 >= 0 && row0 < H, "dd_ground_score: bad arguments");
  DD_REQUIRE((K + B - 1) / B <= GS_MAX_HYP, "dd_ground_score: more than %d hypotheses per image", GS_MAX_HYP);
  cudaStream_t st = (cudaStream_t)stream;
  DD_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int32_t), st));
  const int n = (H - row0) * W;
  int bx = (n + GS_THREADS * 4 - 1) / (GS_THREADS * 4);
  bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
  ground_score_kernel<<<dim3(bx, B), GS_THREADS, 0, st>>>(points, w, B, H, W, row0, K, tol, counts);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}
