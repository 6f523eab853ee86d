// Stand-alone geometry / SSIM layers (tools.py:167-257) for callers outside the fused training step.
#include "warp_photo.cuh"

namespace dd {

constexpr int GE_THREADS = 256;
static inline int ge_blocks(size_t n) { return (int)((n + GE_THREADS - 1) / GE_THREADS < 4736 ? (n + GE_THREADS - 1) / GE_THREADS : 4736); }

__global__ void backproject_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ inv_K, int B, int H, int W,
                                       float* __restrict__ pts) {
  const size_t P = (size_t)H * W, n = (size_t)B * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / P);
    const size_t p = i - (size_t)b * P;
    const float u = (float)(p % W), v = (float)(p / W);
    const float* iK = inv_K + b * 16;
    const float d = __ldg(depth + i);
    float* o = pts + (size_t)b * 4 * P + p;
    o[0] = d * (__ldg(iK + 0) * u + __ldg(iK + 1) * v + __ldg(iK + 2));
    o[P] = d * (__ldg(iK + 4) * u + __ldg(iK + 5) * v + __ldg(iK + 6));
    o[2 * P] = d * (__ldg(iK + 8) * u + __ldg(iK + 9) * v + __ldg(iK + 10));
    o[3 * P] = 1.f;
  }
}

__global__ void backproject_bwd_kernel(const float* __restrict__ gpts, const float* __restrict__ inv_K, int B, int H, int W,
                                       float* __restrict__ gdepth) {
  const size_t P = (size_t)H * W, n = (size_t)B * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / P);
    const size_t p = i - (size_t)b * P;
    const float u = (float)(p % W), v = (float)(p / W);
    const float* iK = inv_K + b * 16;
    const float* g = gpts + (size_t)b * 4 * P + p;
    gdepth[i] = __ldg(g) * (__ldg(iK + 0) * u + __ldg(iK + 1) * v + __ldg(iK + 2)) +
                __ldg(g + P) * (__ldg(iK + 4) * u + __ldg(iK + 5) * v + __ldg(iK + 6)) +
                __ldg(g + 2 * P) * (__ldg(iK + 8) * u + __ldg(iK + 9) * v + __ldg(iK + 10));
  }
}

__global__ void project_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K, const float* __restrict__ T,
                                   int B, int H, int W, float* __restrict__ pix, float* __restrict__ ego) {
  const size_t P = (size_t)H * W, n = (size_t)B * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / P);
    const size_t p = i - (size_t)b * P;
    const float* q = pts + (size_t)b * 4 * P + p;
    const float x = __ldg(q), y = __ldg(q + P), z = __ldg(q + 2 * P), w = __ldg(q + 3 * P);
    Vec4 X = {x, y, z, w};
    if (T) {
      const float* t = T + b * 16;
      X.x = t[0] * x + t[1] * y + t[2] * z + t[3] * w;
      X.y = t[4] * x + t[5] * y + t[6] * z + t[7] * w;
      X.z = t[8] * x + t[9] * y + t[10] * z + t[11] * w;
      X.w = t[12] * x + t[13] * y + t[14] * z + t[15] * w;
    }
    const Proj pr = project_K(K + b * 16, X);
    reinterpret_cast<float2*>(pix)[i] = make_float2(normalise(pr.px, 1.f / (float)(W - 1)), normalise(pr.py, 1.f / (float)(H - 1)));
    float* e = ego + (size_t)b * 3 * P + p;
    e[0] = X.x - x, e[P] = X.y - y, e[2 * P] = X.z - z;
  }
}

// grad_T is accumulated with one atomicAdd per CTA and entry (zero-initialised by the caller)
__global__ void __launch_bounds__(GE_THREADS) project_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                                 const float* __restrict__ T, const float* __restrict__ gpix,
                                                                 const float* __restrict__ gego, int H, int W,
                                                                 float* __restrict__ gpts, float* __restrict__ gT) {
  __shared__ float sh[GE_THREADS / 32][16];
  const size_t P = (size_t)H * W;
  const int b = blockIdx.y;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  const float* Kb = K + b * 16;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float* q = pts + (size_t)b * 4 * P + p;
    const float pv[4] = {__ldg(q), __ldg(q + P), __ldg(q + 2 * P), __ldg(q + 3 * P)};
    Vec4 X = {pv[0], pv[1], pv[2], pv[3]};
    if (T) {
      const float* t = T + b * 16;
      X.x = t[0] * pv[0] + t[1] * pv[1] + t[2] * pv[2] + t[3] * pv[3];
      X.y = t[4] * pv[0] + t[5] * pv[1] + t[6] * pv[2] + t[7] * pv[3];
      X.z = t[8] * pv[0] + t[9] * pv[1] + t[10] * pv[2] + t[11] * pv[3];
      X.w = t[12] * pv[0] + t[13] * pv[1] + t[14] * pv[2] + t[15] * pv[3];
    }
    const Proj pr = project_K(Kb, X);
    float gX[4] = {0.f, 0.f, 0.f, 0.f};
    if (gpix) {
      const float2 g2 = reinterpret_cast<const float2*>(gpix)[(size_t)b * P + p];
      const float gpx = g2.x * 2.f / (float)(W - 1), gpy = g2.y * 2.f / (float)(H - 1);
      const float iz = 1.f / pr.z;
      const float gc0 = gpx * iz, gc1 = gpy * iz, gc2 = -(gpx * pr.px + gpy * pr.py) * iz;
#pragma unroll
      for (int j = 0; j < 4; ++j) gX[j] = Kb[j] * gc0 + Kb[4 + j] * gc1 + Kb[8 + j] * gc2;
    }
    float gp[4] = {0.f, 0.f, 0.f, 0.f};
    if (gego) {
      const float* ge = gego + (size_t)b * 3 * P + p;
      const float e0 = __ldg(ge), e1 = __ldg(ge + P), e2 = __ldg(ge + 2 * P);
      gX[0] += e0, gX[1] += e1, gX[2] += e2;
      gp[0] -= e0, gp[1] -= e1, gp[2] -= e2;
    }
    if (T) {
      const float* t = T + b * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) gp[j] += t[j] * gX[0] + t[4 + j] * gX[1] + t[8 + j] * gX[2] + t[12 + j] * gX[3];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i * 4 + j] += gX[i] * pv[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) gp[j] += gX[j];
    }
    float* go = gpts + (size_t)b * 4 * P + p;
    go[0] = gp[0], go[P] = gp[1], go[2 * P] = gp[2], go[3 * P] = gp[3];
  }
  if (gT && T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float v = 0.f;
      for (int i = 0; i < GE_THREADS / 32; ++i) v += sh[i][threadIdx.x];
      atomicAdd(gT + b * 16 + threadIdx.x, v);
    }
  }
}

struct SsimStats {
  float mu_x, mu_y, sig_x, sig_y, sig_xy;
};

__device__ __forceinline__ SsimStats ssim_stats(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int r,
                                                int c) {
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int rr = reflect1(r + dy, H);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int cc = reflect1(c + dx, W);
      const float xv = __ldg(x + rr * W + cc), yv = __ldg(y + rr * W + cc);
      sx += xv, sy += yv, sxx += xv * xv, syy += yv * yv, sxy += xv * yv;
    }
  }
  SsimStats s;
  s.mu_x = sx / 9.f, s.mu_y = sy / 9.f;
  s.sig_x = sxx / 9.f - s.mu_x * s.mu_x;
  s.sig_y = syy / 9.f - s.mu_y * s.mu_y;
  s.sig_xy = sxy / 9.f - s.mu_x * s.mu_y;
  return s;
}

__global__ void ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int BC, int H, int W,
                                float* __restrict__ out) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const size_t P = (size_t)H * W, n = (size_t)BC * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / P;
    const int p = (int)(i - bc * P);
    const SsimStats s = ssim_stats(x + bc * P, y + bc * P, H, W, p / W, p % W);
    const float nn = (2.f * s.mu_x * s.mu_y + C1) * (2.f * s.sig_xy + C2);
    const float dd_ = (s.mu_x * s.mu_x + s.mu_y * s.mu_y + C1) * (s.sig_x + s.sig_y + C2);
    out[i] = fminf(fmaxf((1.f - nn / dd_) / 2.f, 0.f), 1.f);
  }
}

__global__ void ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ go, int BC,
                                int H, int W, float* __restrict__ gx) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const size_t P = (size_t)H * W, n = (size_t)BC * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / P;
    const int p = (int)(i - bc * P);
    const int r = p / W, c = p % W;
    const float* xp = x + bc * P;
    const float* yp = y + bc * P;
    const float xq = __ldg(xp + p), yq = __ldg(yp + p);
    float g = 0.f;
    for (int dy = -1; dy <= 1; ++dy) {
      const int pr = r + dy;
      if (pr < 0 || pr >= H) continue;
      const float wr = ((r == 1 && pr == 0) || (r == H - 2 && pr == H - 1)) ? 2.f : 1.f;
      for (int dx = -1; dx <= 1; ++dx) {
        const int pc = c + dx;
        if (pc < 0 || pc >= W) continue;
        const float wc = ((c == 1 && pc == 0) || (c == W - 2 && pc == W - 1)) ? 2.f : 1.f;
        const SsimStats s = ssim_stats(xp, yp, H, W, pr, pc);
        const float A1 = 2.f * s.mu_x * s.mu_y + C1, A2 = 2.f * s.sig_xy + C2;
        const float B1 = s.mu_x * s.mu_x + s.mu_y * s.mu_y + C1, B2 = s.sig_x + s.sig_y + C2;
        const float nn = A1 * A2, dn = B1 * B2;
        const float v = (1.f - nn / dn) / 2.f;
        if (!(v >= 0.f && v <= 1.f)) continue;
        const float inv_d = 1.f / dn;
        const float dS_dmu = (2.f * s.mu_y * (A2 - A1) * dn - nn * 2.f * s.mu_x * (B2 - B1)) * inv_d * inv_d;
        const float dS_dxx = -nn * B1 * inv_d * inv_d;
        const float dS_dxy = 2.f * A1 * inv_d;
        const float G = -0.5f * __ldg(go + bc * P + (size_t)pr * W + pc) / 9.f;
        g += wr * wc * G * (dS_dmu + 2.f * xq * dS_dxx + yq * dS_dxy);
      }
    }
    gx[i] = g;
  }
}

}  // namespace dd

extern "C" {
using namespace dd;

int dd_backproject_fwd(const float* depth, const float* inv_K, int B, int H, int W, float* points, void* stream) {
  DD_REQUIRE(depth && inv_K && points && B > 0 && H > 0 && W > 0, "dd_backproject_fwd: bad arguments");
  backproject_fwd_kernel<<<ge_blocks((size_t)B * H * W), GE_THREADS, 0, (cudaStream_t)stream>>>(depth, inv_K, B, H, W, points); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_backproject_bwd(const float* grad_points, const float* inv_K, int B, int H, int W, float* grad_depth, void* stream) {
  DD_REQUIRE(grad_points && inv_K && grad_depth && B > 0 && H > 0 && W > 0, "dd_backproject_bwd: bad arguments");
  backproject_bwd_kernel<<<ge_blocks((size_t)B * H * W), GE_THREADS, 0, (cudaStream_t)stream>>>(grad_points, inv_K, B, H, W, grad_depth); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_project_fwd(const float* points, const float* K, const float* T, int B, int H, int W, float* pix, float* ego,
                   void* stream) {
  DD_REQUIRE(points && K && pix && ego && B > 0 && H > 1 && W > 1, "dd_project_fwd: bad arguments");
  project_fwd_kernel<<<ge_blocks((size_t)B * H * W), GE_THREADS, 0, (cudaStream_t)stream>>>(points, K, T, B, H, W, pix, ego); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_project_bwd(const float* points, const float* K, const float* T, const float* grad_pix, const float* grad_ego, int B,
                   int H, int W, float* grad_points, float* grad_T, void* stream) {
  DD_REQUIRE(points && K && grad_points && B > 0 && H > 1 && W > 1, "dd_project_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_T) DD_CHECK_CUDA(cudaMemsetAsync(grad_T, 0, (size_t)B * 16 * sizeof(float), st));
  const int bx = (int)(((size_t)H * W + GE_THREADS - 1) / GE_THREADS < 148 ? ((size_t)H * W + GE_THREADS - 1) / GE_THREADS : 148);
  project_bwd_kernel<<<dim3(bx, B), GE_THREADS, 0, st>>>(points, K, T, grad_pix, grad_ego, H, W, grad_points, grad_T); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_ssim_fwd(const float* x, const float* y, int BC, int H, int W, float* out, void* stream) {
  DD_REQUIRE(x && y && out && BC > 0 && H >= 2 && W >= 2, "dd_ssim_fwd: bad arguments");
  ssim_fwd_kernel<<<ge_blocks((size_t)BC * H * W), GE_THREADS, 0, (cudaStream_t)stream>>>(x, y, BC, H, W, out); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_ssim_bwd(const float* x, const float* y, const float* grad_out, int BC, int H, int W, float* grad_x, void* stream) {
  DD_REQUIRE(x && y && grad_out && grad_x && BC > 0 && H >= 4 && W >= 4, "dd_ssim_bwd: bad arguments");
  ssim_bwd_kernel<<<ge_blocks((size_t)BC * H * W), GE_THREADS, 0, (cudaStream_t)stream>>>(x, y, grad_out, BC, H, W, grad_x); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
