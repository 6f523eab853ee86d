// Depth-wise dilated 3x3 convolution of the Lite-Mono DilatedConv blocks (reference: networks/depth_encoder.py:148-168
// CDilated = nn.Conv2d(dim, dim, 3, padding=d, dilation=d, groups=dim, bias=False), used at :193/:207) over NCHW fp32
// tensors: forward, data gradient (the same kernel with the taps flipped) and weight gradient.  ATen's depth-wise kernels
// run these 15 layers x 3 encoder passes at ~1/7 of the HBM rate; the op is a pure stream (9 multiply-adds per element), so:
//   thread = 4 consecutive outputs of one row; per tap row it loads the aligned 16-byte vectors that cover the three
//   dilated taps (3 vectors for d <= 4, 5 for d = 6 -- neighbouring threads and rows hit L1) and selects the shifted
//   4-element runs in registers (the dilation is a template parameter, so the selection is static); y written as one
//   16-byte store.  Weight gradient: the same tap loader, 9 running sums per thread, block reduction, per-CTA partials and a
//   fixed-order second stage (deterministic, no atomics).
#include "dd_common.cuh"

namespace dd {

constexpr int DW_THREADS = 256;
constexpr int DW_MAX_CHUNKS = 64;

struct DwArgs {
  const float* x;
  const float* w;     // (C,1,3,3)
  const float* gy;    // weight gradient only
  float* y;
  float* partial;     // weight gradient: [C][chunks][9]
  int B, C, H, W, chunks;
};

// the 4-element run starting at element offset OFF (may be negative) relative to the thread's own aligned vector `q`
// of row `row` (W4 vectors per row); vectors outside [0, W4) are zero padding
template <int OFF>
__device__ __forceinline__ float4 shifted_run(const float4* __restrict__ row, int q, int W4) {
  constexpr int A = OFF >= 0 ? OFF / 4 : -((-OFF + 3) / 4);   // floor(OFF / 4)
  constexpr int R = OFF - 4 * A;                              // 0..3
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const int q0 = q + A;
  const float4 lo = (q0 >= 0 && q0 < W4) ? __ldg(row + q0) : z;
  if (R == 0) return lo;
  const float4 hi = (q0 + 1 >= 0 && q0 + 1 < W4) ? __ldg(row + q0 + 1) : z;
  if (R == 1) return make_float4(lo.y, lo.z, lo.w, hi.x);
  if (R == 2) return make_float4(lo.z, lo.w, hi.x, hi.y);
  return make_float4(lo.w, hi.x, hi.y, hi.z);
}

// taps[kx] = the four inputs under tap column kx (x - D, x, x + D) of one input row
template <int D>
__device__ __forceinline__ void row_taps(const float4* __restrict__ row, int q, int W4, float4 taps[3]) {
  taps[0] = shifted_run<-D>(row, q, W4);
  taps[1] = shifted_run<0>(row, q, W4);
  taps[2] = shifted_run<D>(row, q, W4);
}

template <int D>
__global__ void __launch_bounds__(DW_THREADS) dwconv3_kernel(const __grid_constant__ DwArgs a, int flip) {
  const int W4 = a.W >> 2;
  const long long total = (long long)a.B * a.C * a.H * W4;
  for (long long u = (long long)blockIdx.x * DW_THREADS + threadIdx.x; u < total; u += (long long)gridDim.x * DW_THREADS) {
    const int q = (int)(u % W4);
    const long long t = u / W4;
    const int y = (int)(t % a.H);
    const long long plane = t / a.H;
    const int c = (int)(plane % a.C);
    const float4* xp = reinterpret_cast<const float4*>(a.x) + plane * a.H * W4;
    float wv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wv[k] = __ldg(a.w + c * 9 + (flip ? 8 - k : k));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + (ky - 1) * D;
      if (yy < 0 || yy >= a.H) continue;
      float4 taps[3];
      row_taps<D>(xp + (long long)yy * W4, q, W4, taps);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float wk = wv[ky * 3 + kx];
        acc.x = fmaf(wk, taps[kx].x, acc.x), acc.y = fmaf(wk, taps[kx].y, acc.y);
        acc.z = fmaf(wk, taps[kx].z, acc.z), acc.w = fmaf(wk, taps[kx].w, acc.w);
      }
    }
    reinterpret_cast<float4*>(a.y)[u] = acc;
  }
}

// gw[c][ky][kx] = sum_{b,y,x} gy[b,c,y,x] * x[b,c,y+(ky-1)D,x+(kx-1)D]; grid (chunks, C)
template <int D>
__global__ void __launch_bounds__(DW_THREADS) dwconv3_wgrad_kernel(const __grid_constant__ DwArgs a) {
  __shared__ float red[DW_THREADS / 32][9];
  const int c = blockIdx.y, chunk = blockIdx.x;
  const int W4 = a.W >> 2;
  const int per_img = a.H * W4;
  const long long total = (long long)a.B * per_img;
  float s[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) s[k] = 0.f;
  for (long long u = (long long)chunk * DW_THREADS + threadIdx.x; u < total; u += (long long)a.chunks * DW_THREADS) {
    const int b = (int)(u / per_img), r = (int)(u - (long long)b * per_img);
    const int y = r / W4, q = r - y * W4;
    const long long plane = (long long)b * a.C + c;
    const float4* xp = reinterpret_cast<const float4*>(a.x) + plane * per_img;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gy) + plane * per_img + r);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + (ky - 1) * D;
      if (yy < 0 || yy >= a.H) continue;
      float4 taps[3];
      row_taps<D>(xp + (long long)yy * W4, q, W4, taps);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        s[ky * 3 + kx] += (g.x * taps[kx].x + g.y * taps[kx].y) + (g.z * taps[kx].z + g.w * taps[kx].w);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = warp_sum(s[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
#pragma unroll
    for (int wi = 0; wi < DW_THREADS / 32; ++wi) v += red[wi][threadIdx.x];
    a.partial[((size_t)c * a.chunks + chunk) * 9 + threadIdx.x] = v;
  }
}

__global__ void dwconv3_wgrad_reduce_kernel(const float* __restrict__ partial, int C, int chunks, float* __restrict__ gw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // c * 9 + k
  if (i >= C * 9) return;
  const int c = i / 9, k = i - c * 9;
  double s = 0.0;
  for (int j = 0; j < chunks; ++j) s += (double)partial[((size_t)c * chunks + j) * 9 + k];
  gw[i] = (float)s;
}

static int dw_check(const char* what, int B, int C, int H, int W, int d) {
  DD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "%s: bad shape", what);
  DD_REQUIRE(W % 4 == 0, "%s: W must be a multiple of 4 (got %d)", what, W);
  DD_REQUIRE(d == 1 || d == 2 || d == 3 || d == 4 || d == 6, "%s: dilation %d not built (1, 2, 3, 4, 6)", what, d);
  DD_REQUIRE(C <= 65535, "%s: too many channels", what);
  return DD_OK;
}

#define DD_DW_DISPATCH(D, KERNEL, ...)          \
  do {                                          \
    switch (D) {                                \
      case 1: KERNEL<1> __VA_ARGS__; break;     \
      case 2: KERNEL<2> __VA_ARGS__; break;     \
      case 3: KERNEL<3> __VA_ARGS__; break;     \
      case 4: KERNEL<4> __VA_ARGS__; break;     \
      default: KERNEL<6> __VA_ARGS__; break;    \
    }                                           \
  } while (0)

}  // namespace dd

extern "C" {

size_t dd_dwconv3x3_workspace_bytes(int C) { return C > 0 ? (size_t)C * dd::DW_MAX_CHUNKS * 9 * sizeof(float) : 0; }

int dd_dwconv3x3_fwd(const float* x, const float* w, int B, int C, int H, int W, int dilation, int flip, float* y, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && w && y, "dd_dwconv3x3_fwd: NULL pointer");
  if (int rc = dw_check("dd_dwconv3x3_fwd", B, C, H, W, dilation)) return rc;
  DD_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "dd_dwconv3x3_fwd: x / y must be 16-byte aligned");
  DwArgs a = {};
  a.x = x, a.w = w, a.y = y, a.B = B, a.C = C, a.H = H, a.W = W;
  const long long units = (long long)B * C * H * (W / 4);
  const long long want = (units + DW_THREADS - 1) / DW_THREADS;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  DD_DW_DISPATCH(dilation, dwconv3_kernel, <<<grid, DW_THREADS, 0, st>>>(a, flip));
  count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_dwconv3x3_wgrad(const float* x, const float* grad_y, int B, int C, int H, int W, int dilation, float* grad_w, void* workspace,
                       size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && grad_y && grad_w, "dd_dwconv3x3_wgrad: NULL pointer");
  if (int rc = dw_check("dd_dwconv3x3_wgrad", B, C, H, W, dilation)) return rc;
  DD_REQUIRE((((uintptr_t)x | (uintptr_t)grad_y) & 15) == 0, "dd_dwconv3x3_wgrad: x / grad_y must be 16-byte aligned");
  if (!workspace || workspace_bytes < dd_dwconv3x3_workspace_bytes(C)) {
    set_error("dd_dwconv3x3_wgrad: workspace too small (%zu < %zu)", workspace_bytes, dd_dwconv3x3_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  DwArgs a = {};
  a.x = x, a.gy = grad_y, a.partial = reinterpret_cast<float*>(workspace), a.B = B, a.C = C, a.H = H, a.W = W;
  const long long units = (long long)B * H * (W / 4);
  long long k = (148 * 8 + C - 1) / C;
  const long long cap = (units + DW_THREADS * 2 - 1) / (DW_THREADS * 2);
  k = k > cap ? cap : k;
  k = k > DW_MAX_CHUNKS ? DW_MAX_CHUNKS : (k < 1 ? 1 : k);
  a.chunks = (int)k;
  cudaStream_t st = (cudaStream_t)stream;
  DD_DW_DISPATCH(dilation, dwconv3_wgrad_kernel, <<<dim3(a.chunks, C), DW_THREADS, 0, st>>>(a));
  dwconv3_wgrad_reduce_kernel<<<(C * 9 + 127) / 128, 128, 0, st>>>(a.partial, C, a.chunks, grad_w);
  count_launches(2);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
