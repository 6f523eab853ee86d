// channels_last LayerNorm of the Lite-Mono LGFI blocks (reference: networks/depth_encoder.py:90-104 LayerNorm.forward ->
// F.layer_norm over the last dimension, used at :261 norm_xca and :266 norm) for rows of C = 64 / 128 / 224 floats.
// ATen's kernel gives one CTA to every row; with 256-byte rows that is one warp-load per CTA and 0.35 ms for the 63 MB
// stage-1 map.  Here a row belongs to LANES lanes of a warp (16 lanes = two rows per warp for C = 64), every lane keeps its
// PER float4 of the row in registers, mean and variance are two shuffle reductions over registers (two-pass: no
// cancellation), and a persistent grid walks the rows.  HBM-bound: x read once, y written once (+ 8 bytes of statistics per
// row); backward reads x and grad_y once, writes grad_x once and reduces grad_gamma / grad_beta through per-CTA partials
// and a fixed-order second stage (deterministic, no atomics).
#include "dd_common.cuh"

namespace dd {

constexpr int LN_THREADS = 256;

struct LnArgs {
  const float* x;
  const float* gy;
  const float* gamma;
  const float* beta;
  float* y;        // forward: y, backward: grad_x
  float* mean;
  float* rstd;
  float* partial;  // backward: [grid][2][C]
  long long M;
  int C;
  float eps;
};

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LANES, int PER>
__global__ void __launch_bounds__(LN_THREADS) layernorm_fwd_kernel(const __grid_constant__ LnArgs a) {
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31, sub = lane % LANES, grp = lane / LANES;
  const int C4 = a.C >> 2;
  const long long warp_global = (long long)blockIdx.x * (LN_THREADS / 32) + (threadIdx.x >> 5);
  const long long warps_total = (long long)gridDim.x * (LN_THREADS / 32);
  float4 w[PER], bsh[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = sub + LANES * j;
    w[j] = (a.gamma && i < C4) ? __ldg(reinterpret_cast<const float4*>(a.gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    bsh[j] = (a.beta && i < C4) ? __ldg(reinterpret_cast<const float4*>(a.beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.f / (float)a.C;
  for (long long row = warp_global * ROWS_PER_WARP + grp; row < a.M + grp; row += warps_total * ROWS_PER_WARP) {
    const bool live = row < a.M;   // (whole warp stays in the loop for the shuffles)
    const float4* xr = reinterpret_cast<const float4*>(a.x) + row * C4;
    float4 v[PER];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = sub + LANES * j;
      v[j] = (live && i < C4) ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mu = group_sum<LANES>(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = sub + LANES * j;
      if (i < C4) {
        const float d0 = v[j].x - mu, d1 = v[j].y - mu, d2 = v[j].z - mu, d3 = v[j].w - mu;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    const float rs = rsqrtf(group_sum<LANES>(q) * inv_c + a.eps);
    if (!live) continue;
    float4* yr = reinterpret_cast<float4*>(a.y) + row * C4;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = sub + LANES * j;
      if (i < C4)
        yr[i] = make_float4(fmaf((v[j].x - mu) * rs, w[j].x, bsh[j].x), fmaf((v[j].y - mu) * rs, w[j].y, bsh[j].y),
                            fmaf((v[j].z - mu) * rs, w[j].z, bsh[j].z), fmaf((v[j].w - mu) * rs, w[j].w, bsh[j].w));
    }
    if (sub == 0) a.mean[row] = mu, a.rstd[row] = rs;
  }
}

template <int LANES, int PER>
__global__ void __launch_bounds__(LN_THREADS) layernorm_bwd_kernel(const __grid_constant__ LnArgs a) {
  constexpr int ROWS_PER_WARP = 32 / LANES;
  extern __shared__ float red[];   // [warps][ROWS_PER_WARP][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LANES, grp = lane / LANES;
  const int C4 = a.C >> 2;
  const long long warp_global = (long long)blockIdx.x * (LN_THREADS / 32) + warp;
  const long long warps_total = (long long)gridDim.x * (LN_THREADS / 32);
  float4 w[PER], dw[PER], db[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = sub + LANES * j;
    w[j] = (a.gamma && i < C4) ? __ldg(reinterpret_cast<const float4*>(a.gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    dw[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.f / (float)a.C;
  for (long long row = warp_global * ROWS_PER_WARP + grp; row < a.M + grp; row += warps_total * ROWS_PER_WARP) {
    const bool live = row < a.M;
    const float4* xr = reinterpret_cast<const float4*>(a.x) + row * C4;
    const float4* gr = reinterpret_cast<const float4*>(a.gy) + row * C4;
    const float mu = live ? __ldg(a.mean + row) : 0.f, rs = live ? __ldg(a.rstd + row) : 0.f;
    float4 xh[PER], g[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = sub + LANES * j;
      const bool ok = live && i < C4;
      const float4 xv = ok ? __ldg(xr + i) : make_float4(mu, mu, mu, mu);
      g[j] = ok ? __ldg(gr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dw[j].x = fmaf(g[j].x, xh[j].x, dw[j].x), dw[j].y = fmaf(g[j].y, xh[j].y, dw[j].y);
      dw[j].z = fmaf(g[j].z, xh[j].z, dw[j].z), dw[j].w = fmaf(g[j].w, xh[j].w, dw[j].w);
      db[j].x += g[j].x, db[j].y += g[j].y, db[j].z += g[j].z, db[j].w += g[j].w;
      g[j] = make_float4(g[j].x * w[j].x, g[j].y * w[j].y, g[j].z * w[j].z, g[j].w * w[j].w);   // gradient w.r.t. xhat
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
    const float m1 = group_sum<LANES>(s1) * inv_c, m2 = group_sum<LANES>(s2) * inv_c;
    if (!live || a.y == nullptr) continue;
    float4* dr = reinterpret_cast<float4*>(a.y) + row * C4;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = sub + LANES * j;
      if (i < C4)
        dr[i] = make_float4(rs * ((g[j].x - m1) - xh[j].x * m2), rs * ((g[j].y - m1) - xh[j].y * m2),
                            rs * ((g[j].z - m1) - xh[j].z * m2), rs * ((g[j].w - m1) - xh[j].w * m2));
    }
  }
  if (a.partial == nullptr) return;
  // per-CTA column sums: every (warp, row group) deposits its registers, then thread c adds them in a fixed order
  float* mine = red + (size_t)(warp * ROWS_PER_WARP + grp) * 2 * a.C;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = sub + LANES * j;
    if (i < C4) {
      reinterpret_cast<float4*>(mine)[i] = dw[j];
      reinterpret_cast<float4*>(mine + a.C)[i] = db[j];
    }
  }
  __syncthreads();
  constexpr int SLOTS = (LN_THREADS / 32) * ROWS_PER_WARP;
  for (int c = threadIdx.x; c < 2 * a.C; c += LN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) s += red[(size_t)k * 2 * a.C + c];
    a.partial[(size_t)blockIdx.x * 2 * a.C + c] = s;
  }
}

// grad_gamma[c] = sum over CTAs of partial[cta][0][c], grad_beta likewise: CTA = 32 columns x 32 row groups (1024 threads), four
// independent loads per step (one load per step over 8 groups = 148 dependent L2 round trips = 63 us, ncu); fixed-order double sums
__global__ void __launch_bounds__(1024) layernorm_reduce_kernel(const float* __restrict__ partial, int ctas, int C, float* __restrict__ ggamma,
                                                                float* __restrict__ gbeta) {
  __shared__ double sh[32][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
  double s = 0.0;
  if (c < 2 * C) {
    for (int k = grp; k < ctas; k += 4 * 32) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (k + 32 * j < ctas) ? partial[(size_t)(k + 32 * j) * 2 * C + c] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (double)v[j];
    }
  }
  sh[grp][threadIdx.x & 31] = s;
  __syncthreads();
  if (grp == 0 && c < 2 * C) {
#pragma unroll 8
    for (int g = 1; g < 32; ++g) s += sh[g][threadIdx.x];
    if (c < C) { if (ggamma) ggamma[c] = (float)s; }
    else if (gbeta) gbeta[c - C] = (float)s;
  }
}

static int ln_grid(long long M, int rows_per_warp) {
  const long long rows_per_cta = (long long)(LN_THREADS / 32) * rows_per_warp;
  long long g = (M + rows_per_cta - 1) / rows_per_cta;
  const long long cap = 148 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int ln_check(const char* what, long long M, int C) {
  DD_REQUIRE(M > 0 && C > 0, "%s: bad shape M=%lld C=%d", what, M, C);
  DD_REQUIRE(C % 4 == 0 && C <= 512, "%s: C must be a multiple of 4 and <= 512 (got %d)", what, C);
  return DD_OK;
}

#define DD_LN_DISPATCH(KERNEL, ...)                                    \
  do {                                                                 \
    if (C4 <= 16) KERNEL<16, 1> __VA_ARGS__;                           \
    else if (C4 <= 32) KERNEL<32, 1> __VA_ARGS__;                      \
    else if (C4 <= 64) KERNEL<32, 2> __VA_ARGS__;                      \
    else KERNEL<32, 4> __VA_ARGS__;                                    \
  } while (0)

}  // namespace dd

extern "C" {

size_t dd_layernorm_workspace_bytes(int C) { return C > 0 ? (size_t)148 * 8 * 2 * C * sizeof(float) : 0; }

int dd_layernorm_fwd(const float* x, long long M, int C, const float* gamma, const float* beta, float eps, float* y, float* mean,
                     float* rstd, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && y && mean && rstd, "dd_layernorm_fwd: NULL pointer");
  if (int rc = ln_check("dd_layernorm_fwd", M, C)) return rc;
  DD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && (!gamma || ((uintptr_t)gamma & 15) == 0) && (!beta || ((uintptr_t)beta & 15) == 0),
             "dd_layernorm_fwd: pointers must be 16-byte aligned");
  LnArgs a = {};
  a.x = x, a.gamma = gamma, a.beta = beta, a.y = y, a.mean = mean, a.rstd = rstd, a.M = M, a.C = C, a.eps = eps;
  const int C4 = C / 4;
  const int grid = ln_grid(M, C4 <= 16 ? 2 : 1);
  cudaStream_t st = (cudaStream_t)stream;
  DD_LN_DISPATCH(layernorm_fwd_kernel, <<<grid, LN_THREADS, 0, st>>>(a));
  count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_layernorm_bwd(const float* x, const float* grad_y, long long M, int C, const float* gamma, const float* mean, const float* rstd,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && grad_y && mean && rstd, "dd_layernorm_bwd: NULL pointer");
  DD_REQUIRE(grad_x || grad_gamma || grad_beta, "dd_layernorm_bwd: no gradient requested");
  if (int rc = ln_check("dd_layernorm_bwd", M, C)) return rc;
  DD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)grad_y & 15) == 0 && (!grad_x || ((uintptr_t)grad_x & 15) == 0) &&
                 (!gamma || ((uintptr_t)gamma & 15) == 0), "dd_layernorm_bwd: pointers must be 16-byte aligned");
  const bool affine = grad_gamma || grad_beta;
  if (affine && (!workspace || workspace_bytes < dd_layernorm_workspace_bytes(C))) {
    set_error("dd_layernorm_bwd: workspace too small (%zu < %zu)", workspace_bytes, dd_layernorm_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  LnArgs a = {};
  a.x = x, a.gy = grad_y, a.gamma = gamma, a.y = grad_x, a.mean = const_cast<float*>(mean), a.rstd = const_cast<float*>(rstd), a.M = M, a.C = C;
  a.partial = affine ? reinterpret_cast<float*>(workspace) : nullptr;
  const int C4 = C / 4;
  const int rpw = C4 <= 16 ? 2 : 1;
  const int grid = ln_grid(M, rpw);
  const size_t smem = affine ? (size_t)(LN_THREADS / 32) * rpw * 2 * C * sizeof(float) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  DD_LN_DISPATCH(layernorm_bwd_kernel, <<<grid, LN_THREADS, smem, st>>>(a));
  count_launches(1);
  if (affine) {
    layernorm_reduce_kernel<<<(2 * C + 31) / 32, 1024, 0, st>>>(a.partial, grid, C, grad_gamma, grad_beta);
    count_launches(1);
  }
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
