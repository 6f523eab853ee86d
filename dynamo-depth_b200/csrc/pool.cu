// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the ResNet trunks (reference: networks/resnet_encoder.py:18,:130) on
// channels_last (N,H,W,C) fp32 activations.  ATen's NHWC backward scatters with atomics into a zeroed gradient (0.5 ms for
// the 252 MB map at bs32); here the forward pass records the position of the maximum inside its window (one byte per
// output: first maximum in row-major scan order, NaN propagates -- ATen's rule) and the backward pass GATHERS: every input
// element sums the gradients of the one, two or four windows that cover it and selected it.  Both passes are streams of
// 16-byte channel vectors: the input map is read once / written once, no atomics, no memset.
#include "dd_common.cuh"

namespace dd {

constexpr int MP_THREADS = 256;

__device__ __forceinline__ void take_max(float v, int code, float& best, int& arg) {
  if (v > best || v != v) best = v, arg = code;
}

__global__ void __launch_bounds__(MP_THREADS) maxpool_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, uchar4* __restrict__ idx,
                                                                 int B, int H, int W, int C4, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long u = (long long)blockIdx.x * MP_THREADS + threadIdx.x; u < total; u += (long long)gridDim.x * MP_THREADS) {
    const int c = (int)(u % C4);
    long long t = u / C4;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), b = (int)(t / Ho);
    const float ninf = -__int_as_float(0x7f800000);
    float4 best = make_float4(ninf, ninf, ninf, ninf);
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    bool first = true;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const float4 v = __ldg(x + (((long long)b * H + iy) * W + ix) * C4 + c);
        const int code = ky * 3 + kx;
        if (first) {   // ATen starts from the first in-range element's index (maxval = -inf, comparison below still applies)
          a0 = a1 = a2 = a3 = code;
          first = false;
        }
        take_max(v.x, code, best.x, a0), take_max(v.y, code, best.y, a1), take_max(v.z, code, best.z, a2), take_max(v.w, code, best.w, a3);
      }
    }
    y[u] = best;
    idx[u] = make_uchar4((unsigned char)a0, (unsigned char)a1, (unsigned char)a2, (unsigned char)a3);
  }
}

__global__ void __launch_bounds__(MP_THREADS) maxpool_bwd_kernel(const float4* __restrict__ gy, const uchar4* __restrict__ idx, float4* __restrict__ gx,
                                                                 int B, int H, int W, int C4, int Ho, int Wo) {
  const long long total = (long long)B * H * W * C4;
  for (long long u = (long long)blockIdx.x * MP_THREADS + threadIdx.x; u < total; u += (long long)gridDim.x * MP_THREADS) {
    const int c = (int)(u % C4);
    long long t = u / C4;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H), b = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // windows covering (iy, ix): oy with 2 oy - 1 <= iy <= 2 oy + 1, i.e. iy >> 1 and (iy + 1) >> 1 (the same window for even
    // iy); all candidate loads are issued before any is used
    const int oy0 = iy >> 1, oy1 = (iy + 1) >> 1, ox0 = ix >> 1, ox1 = (ix + 1) >> 1;
    const int oys[2] = {oy0, oy1}, oxs[2] = {ox0, ox1};
    uchar4 sel[4];
    float4 g[4];
    int code[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int oy = oys[k >> 1], ox = oxs[k & 1];
      const bool ok = oy < Ho && ox < Wo && !((k >> 1) == 1 && oy1 == oy0) && !((k & 1) == 1 && ox1 == ox0);
      code[k] = ok ? (iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)) : 255;
      const long long o = (((long long)b * Ho + min(oy, Ho - 1)) * Wo + min(ox, Wo - 1)) * C4 + c;
      sel[k] = __ldg(idx + o);
      g[k] = __ldg(gy + o);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (sel[k].x == code[k]) acc.x += g[k].x;
      if (sel[k].y == code[k]) acc.y += g[k].y;
      if (sel[k].z == code[k]) acc.z += g[k].z;
      if (sel[k].w == code[k]) acc.w += g[k].w;
    }
    gx[u] = acc;
  }
}

static int mp_check(const char* what, int B, int H, int W, int C) {
  DD_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", what);
  DD_REQUIRE(C % 4 == 0, "%s: C must be a multiple of 4 (got %d)", what, C);
  return DD_OK;
}

static int mp_grid(long long units) {
  const long long want = (units + MP_THREADS - 1) / MP_THREADS;
  return (int)(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
}

}  // namespace dd

extern "C" {

int dd_maxpool3x3s2_nhwc_fwd(const float* x, int B, int H, int W, int C, float* y, unsigned char* argmax, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && y && argmax, "dd_maxpool3x3s2_nhwc_fwd: NULL pointer");
  if (int rc = mp_check("dd_maxpool3x3s2_nhwc_fwd", B, H, W, C)) return rc;
  DD_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0 && ((uintptr_t)argmax & 3) == 0, "dd_maxpool3x3s2_nhwc_fwd: misaligned pointer");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  maxpool_fwd_kernel<<<mp_grid((long long)B * Ho * Wo * (C / 4)), MP_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), reinterpret_cast<uchar4*>(argmax), B, H, W, C / 4, Ho, Wo);
  count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_maxpool3x3s2_nhwc_bwd(const float* grad_y, const unsigned char* argmax, int B, int H, int W, int C, float* grad_x, void* stream) {
  using namespace dd;
  DD_REQUIRE(grad_y && argmax && grad_x, "dd_maxpool3x3s2_nhwc_bwd: NULL pointer");
  if (int rc = mp_check("dd_maxpool3x3s2_nhwc_bwd", B, H, W, C)) return rc;
  DD_REQUIRE((((uintptr_t)grad_y | (uintptr_t)grad_x) & 15) == 0 && ((uintptr_t)argmax & 3) == 0, "dd_maxpool3x3s2_nhwc_bwd: misaligned pointer");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  maxpool_bwd_kernel<<<mp_grid((long long)B * H * W * (C / 4)), MP_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(grad_y), reinterpret_cast<const uchar4*>(argmax), reinterpret_cast<float4*>(grad_x), B, H, W, C / 4, Ho, Wo);
  count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
