// Input colour pyramid: ('color', 0, s) = clamp(Resize((H/2^s, W/2^s), BICUBIC, antialias=True)(('color', 0, s-1)), 0, 1)
// (Trainer.py:80, 729-734).  torchvision's tensor path is F.interpolate(mode='bicubic', antialias=True,
// align_corners=False), i.e. ATen's _upsample_bicubic2d_aa: per output index i, centre = scale*(i+0.5),
// support = 2*scale, taps [int(centre-support+0.5), int(centre+support+0.5)) clipped to the axis, weights
// cubic_{a=-0.5}((j - centre + 0.5)/scale) normalised by their sum; the 2-D result is the tensor product.
// For the exact x1/2 step used here that is 8 taps per axis (-3,-9,29,111,111,29,-9,-3)/256 in the interior and
// renormalised truncated windows at the borders.
//
// One CTA produces a 16x32 output tile: the 38x70 input patch is staged in shared memory, filtered along x into a
// 38x32 intermediate, then along y; every input element is read from HBM once (plus the 6-pixel tile halo).
#include "dd_common.cuh"

namespace dd {

constexpr int PY_TH = 16, PY_TW = 32;            // output tile
constexpr int PY_IH = 2 * PY_TH + 6, PY_IW = 2 * PY_TW + 6;   // 38 x 70 input patch
constexpr int PY_IP = PY_IW + 1;                 // 71
constexpr int PY_THREADS = 256;

__device__ __forceinline__ float cubic_aa(float t) {   // ATen bicubic_filter, a = -0.5
  const float a = -0.5f;
  t = fabsf(t);
  if (t < 1.f) return ((a + 2.f) * t - (a + 3.f)) * t * t + 1.f;
  if (t < 2.f) return (((t - 5.f) * t + 8.f) * t - 4.f) * a;
  return 0.f;
}

// taps of output index i on an axis of n_in = 2*n_out samples: first input index and up to 8 normalised weights
__device__ __forceinline__ void aa_taps(int i, int n_in, int& first, int& count, float (&w)[8]) {
  const float centre = 2.f * ((float)i + 0.5f);
  first = max(0, (int)(centre - 4.f + 0.5f));
  const int last = min(n_in, (int)(centre + 4.f + 0.5f));
  count = last - first;
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w[j] = j < count ? cubic_aa(((float)(j + first) - centre + 0.5f) * 0.5f) : 0.f;
    tot += w[j];
  }
  const float inv = 1.f / tot;
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] *= inv;
}

__global__ void __launch_bounds__(PY_THREADS) pyramid_half_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W) {
  __shared__ float in_s[PY_IH * PY_IP];
  __shared__ float mid_s[PY_IH * (PY_TW + 1)];
  const int ho = H >> 1, wo = W >> 1;
  const int oy0 = blockIdx.y * PY_TH, ox0 = blockIdx.x * PY_TW;
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;   // interior windows start 3 samples before 2*i
  const float* xp = x + (size_t)blockIdx.z * H * W;
  float* op = out + (size_t)blockIdx.z * ho * wo;
  const int tid = threadIdx.x;

  for (int i = tid; i < PY_IH * PY_IW; i += PY_THREADS) {
    const int r = i / PY_IW, c = i - r * PY_IW;
    const int gy = iy0 + r, gx = ix0 + c;
    in_s[r * PY_IP + c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // x pass: thread = (patch row, output column)
  for (int i = tid; i < PY_IH * PY_TW; i += PY_THREADS) {
    const int r = i / PY_TW, c = i - r * PY_TW;
    const int ox = ox0 + c;
    float v = 0.f;
    if (ox < wo) {
      int first, count;
      float w[8];
      aa_taps(ox, W, first, count, w);
      const float* row = in_s + r * PY_IP + (first - ix0);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < count) v += w[j] * row[j];
    }
    mid_s[r * (PY_TW + 1) + c] = v;
  }
  __syncthreads();
  // y pass + clamp
  for (int i = tid; i < PY_TH * PY_TW; i += PY_THREADS) {
    const int r = i / PY_TW, c = i - r * PY_TW;
    const int oy = oy0 + r, ox = ox0 + c;
    if (oy >= ho || ox >= wo) continue;
    int first, count;
    float w[8];
    aa_taps(oy, H, first, count, w);
    const float* col = mid_s + (first - iy0) * (PY_TW + 1) + c;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < count) v += w[j] * col[j * (PY_TW + 1)];
    op[(size_t)oy * wo + ox] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

}  // namespace dd

extern "C" int dd_pyramid_half_fwd(const float* x, int BC, int H, int W, float* out, void* stream) {
  DD_REQUIRE(x != nullptr && out != nullptr, "dd_pyramid_half_fwd: NULL tensor");
  DD_REQUIRE(BC > 0 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0, "dd_pyramid_half_fwd: H=%d, W=%d must be even and >= 2", H, W);
  DD_REQUIRE(BC <= 65535, "dd_pyramid_half_fwd: B*C=%d exceeds the grid limit", BC);
  const dim3 grid((W / 2 + dd::PY_TW - 1) / dd::PY_TW, (H / 2 + dd::PY_TH - 1) / dd::PY_TH, BC);
  dd::pyramid_half_kernel<<<grid, dd::PY_THREADS, 0, (cudaStream_t)stream>>>(x, out, H, W);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}
