// Cross-covariance attention core of the Lite-Mono LGFI blocks (reference: networks/depth_encoder.py:63-83 XCA.forward,
// everything between the qkv and proj linear layers):
//   q, k, v = qkv (B,N,3,heads,d) -> (B,heads,d,N);  q, k L2-normalised over the N tokens;
//   attn = softmax_j( (q @ k^T)[i][j] * temperature[head] );  out (B,N,C) = (attn @ v) back in token-major order.
// PyTorch runs this as permuted views + strided element-wise kernels + two skinny batched GEMMs (reduction length N, 8 x 8
// to 28 x 28 outputs) + copies: 0.7 ms forward / 2.1 ms backward for the 48 x 160 stage at bs32.  The arithmetic is tiny
// (C * d multiply-adds per token); the op is a stream over the token-major (B,N,3C) tensor.  Here:
//   forward   gram     per (image, token chunk): partial d x d Gram rows  sum_n q_i[n] k_j[n]  and the squared norms
//             softmax  per image: reduce the partials, S = G / (|q_i| |k_j|), A = softmax(S * temperature)   [tiny]
//             apply    out[n][(h,i)] = sum_j A[i][j] v_j[n]
//   backward  gram     gA[i][j] = sum_n gout_i[n] v_j[n]
//             softmax  backward of softmax / temperature / normalisation -> coefficient matrices                  [tiny]
//             apply    gq_i = sum_j Mq[i][j] k_j - cq_i q_i,  gk_j = sum_i Mq[i][j] q_i - ck_j k_j,  gv_j = sum_i A[i][j] gout_i
// One thread per channel c = (head, i); a tile of 16 tokens is staged in shared memory by coalesced row loads and every
// thread reads its own element plus the d elements of its head (broadcast).  Reductions over tokens go through per-CTA
// partials and a fixed-order second stage: deterministic, no atomics.  qkv is read twice forward and 1 1/3 times backward.
// Round 2b: the token-streaming kernels (gram, apply, backward apply) fetch their
// 16-token x C tiles with tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) into a two-stage ring: one elected thread issues
// the box copies of tile i+1 while all threads compute on tile i (mbarrier complete_tx hand-over), so no thread spends issue
// slots or registers on the staging loads.  DD_NO_TMA=1 (or a failed cuTensorMapEncodeTiled) selects the thread-staged kernels.
#include <cuda.h>   // CUtensorMap + enums only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked
#include <stdlib.h>

#include "tc_common.cuh"

namespace dd {

constexpr int XT = 16;   // tokens per shared-memory tile

struct XcaTok {
  int N, C, tokens_per_cta, chunks;
};

// partial[((b * chunks + chunk) * C + c) * (D + 2) + {0..D-1: sum_n a_c[n] b_{h0+j}[n], D: sum a_c^2, D+1: sum b_c^2}]
template <int D>
__global__ void xca_gram_kernel(const float* __restrict__ a_src, long long a_stride, const float* __restrict__ b_src, long long b_stride,
                                XcaTok g, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* as = sm;
  float* bs = sm + XT * g.C;
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  float G[D];
#pragma unroll
  for (int j = 0; j < D; ++j) G[j] = 0.f;
  float na = 0.f, nb = 0.f;
  for (int t0 = n0; t0 < n1; t0 += XT) {
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < XT; ++t) {
      const int n = t0 + t;
      const bool ok = n < n1;
      as[t * g.C + c] = ok ? __ldg(a_src + ((long long)b * g.N + n) * a_stride + c) : 0.f;
      bs[t * g.C + c] = ok ? __ldg(b_src + ((long long)b * g.N + n) * b_stride + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int t = 0; t < XT; ++t) {
      const float ai = as[t * g.C + c], bc = bs[t * g.C + c];
      na = fmaf(ai, ai, na), nb = fmaf(bc, bc, nb);
      const float* bh = bs + t * g.C + h0;
#pragma unroll
      for (int j = 0; j < D; ++j) G[j] = fmaf(ai, bh[j], G[j]);
    }
  }
  float* p = partial + (((size_t)b * g.chunks + blockIdx.x) * g.C + c) * (D + 2);
#pragma unroll
  for (int j = 0; j < D; ++j) p[j] = G[j];
  p[D] = na, p[D + 1] = nb;
}

// dst[b][n][c] = sum_j M[b][c][j] * src[b][n][h0 + j]
template <int D>
__global__ void xca_apply_kernel(const float* __restrict__ M, const float* __restrict__ src, long long src_stride, float* __restrict__ dst,
                                 long long dst_stride, XcaTok g) {
  extern __shared__ float sm[];
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  float m[D];
#pragma unroll
  for (int j = 0; j < D; ++j) m[j] = __ldg(M + ((size_t)b * g.C + c) * D + j);
  for (int t0 = n0; t0 < n1; t0 += XT) {
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < XT; ++t) {
      const int n = t0 + t;
      sm[t * g.C + c] = n < n1 ? __ldg(src + ((long long)b * g.N + n) * src_stride + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int t = 0; t < XT; ++t) {
      const int n = t0 + t;
      if (n >= n1) break;
      const float* sh = sm + t * g.C + h0;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(m[j], sh[j], s);
      dst[((long long)b * g.N + n) * dst_stride + c] = s;
    }
  }
}

// forward statistics of one image: A, S (B,C,D), rq, rk (B,C)
template <int D>
__global__ void xca_softmax_kernel(const float* __restrict__ partial, int chunks, int C, const float* __restrict__ temp, float* __restrict__ A,
                                   float* __restrict__ S, float* __restrict__ rq, float* __restrict__ rk) {
  extern __shared__ float rk_s[];
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.x;
  double G[D], na = 0.0, nb = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) G[j] = 0.0;
#pragma unroll(D > 16 ? 2 : 4)   // several chunks' loads in flight (one chunk per L2 round trip made this tiny kernel 21 us)
  for (int k = 0; k < chunks; ++k) {
    const float* p = partial + (((size_t)b * chunks + k) * C + c) * (D + 2);
    float v[D + 2];
#pragma unroll
    for (int j = 0; j < D + 2; ++j) v[j] = p[j];
#pragma unroll
    for (int j = 0; j < D; ++j) G[j] += (double)v[j];
    na += (double)v[D], nb += (double)v[D + 1];
  }
  // F.normalize(dim=-1): x / max(||x||, 1e-12)
  const float rqc = 1.f / fmaxf((float)sqrt(na), 1e-12f), rkc = 1.f / fmaxf((float)sqrt(nb), 1e-12f);
  rk_s[c] = rkc;
  __syncthreads();
  const float tp = __ldg(temp + c / D);
  float s[D], mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    s[j] = (float)G[j] * rqc * rk_s[h0 + j];
    mx = fmaxf(mx, s[j] * tp);
  }
  float e[D], sum = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) e[j] = expf(s[j] * tp - mx), sum += e[j];
  const float inv = 1.f / sum;
  float* Ap = A + ((size_t)b * C + c) * D;
  float* Sp = S + ((size_t)b * C + c) * D;
#pragma unroll
  for (int j = 0; j < D; ++j) Ap[j] = e[j] * inv, Sp[j] = s[j];
  rq[(size_t)b * C + c] = rqc, rk[(size_t)b * C + c] = rkc;
}

// backward statistics of one image -> Mq, MqT, AT (B,C,D), cq, ck, gtemp_part (B,C)
template <int D>
__global__ void xca_softmax_bwd_kernel(const float* __restrict__ partial, int chunks, int C, const float* __restrict__ temp,
                                       const float* __restrict__ A, const float* __restrict__ S, const float* __restrict__ rq,
                                       const float* __restrict__ rk, float* __restrict__ Mq, float* __restrict__ MqT, float* __restrict__ AT,
                                       float* __restrict__ cq, float* __restrict__ ck, float* __restrict__ gtemp_part) {
  extern __shared__ float sm[];   // rk_s[C] | buf[C][D + 1]
  float* rk_s = sm;
  float* buf = sm + C;
  const int c = threadIdx.x, h0 = (c / D) * D, jj = c - h0, b = blockIdx.x;
  double gAd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) gAd[j] = 0.0;
#pragma unroll(D > 16 ? 2 : 4)
  for (int k = 0; k < chunks; ++k) {
    const float* p = partial + (((size_t)b * chunks + k) * C + c) * (D + 2);
    float v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = p[j];
#pragma unroll
    for (int j = 0; j < D; ++j) gAd[j] += (double)v[j];
  }
  const float rqc = __ldg(rq + (size_t)b * C + c), rkc = __ldg(rk + (size_t)b * C + c);
  rk_s[c] = rkc;
  const float tp = __ldg(temp + c / D);
  float a[D], s[D], gS[D];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    a[j] = __ldg(A + ((size_t)b * C + c) * D + j), s[j] = __ldg(S + ((size_t)b * C + c) * D + j);
    dot = fmaf(a[j], (float)gAd[j], dot);
  }
  float gt = 0.f, rowdot = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const float gsp = a[j] * ((float)gAd[j] - dot);   // gradient w.r.t. S * temperature
    gt = fmaf(gsp, s[j], gt);
    gS[j] = gsp * tp;
    rowdot = fmaf(gS[j], s[j], rowdot);
  }
  gtemp_part[(size_t)b * C + c] = gt;
  cq[(size_t)b * C + c] = rqc * rqc * rowdot;
  __syncthreads();   // rk_s
  float mq[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    mq[j] = rqc * gS[j] * rk_s[h0 + j];
    Mq[((size_t)b * C + c) * D + j] = mq[j];
    buf[c * (D + 1) + j] = mq[j];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < D; ++i) MqT[((size_t)b * C + c) * D + i] = buf[(h0 + i) * (D + 1) + jj];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < D; ++j) buf[c * (D + 1) + j] = gS[j] * s[j];
  __syncthreads();
  float col = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) col += buf[(h0 + i) * (D + 1) + jj];
  ck[(size_t)b * C + c] = rkc * rkc * col;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < D; ++j) buf[c * (D + 1) + j] = a[j];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < D; ++i) AT[((size_t)b * C + c) * D + i] = buf[(h0 + i) * (D + 1) + jj];
}

// grad_qkv[b][n][{q, k, v}][c] from q, k, v, grad_out
template <int D>
__global__ void xca_bwd_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ gout, const float* __restrict__ Mq,
                                     const float* __restrict__ MqT, const float* __restrict__ AT, const float* __restrict__ cq,
                                     const float* __restrict__ ck, float* __restrict__ gqkv, XcaTok g) {
  extern __shared__ float sm[];
  float* qs = sm;
  float* ks = sm + XT * g.C;
  float* gs = sm + 2 * XT * g.C;
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y, C = g.C;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  float mq[D], mqt[D], at[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    mq[j] = __ldg(Mq + ((size_t)b * C + c) * D + j), mqt[j] = __ldg(MqT + ((size_t)b * C + c) * D + j);
    at[j] = __ldg(AT + ((size_t)b * C + c) * D + j);
  }
  const float cqc = __ldg(cq + (size_t)b * C + c), ckc = __ldg(ck + (size_t)b * C + c);
  for (int t0 = n0; t0 < n1; t0 += XT) {
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < XT; ++t) {
      const int n = t0 + t;
      const bool ok = n < n1;
      const float* row = qkv + ((long long)b * g.N + n) * 3 * C;
      qs[t * C + c] = ok ? __ldg(row + c) : 0.f;
      ks[t * C + c] = ok ? __ldg(row + C + c) : 0.f;
      gs[t * C + c] = ok ? __ldg(gout + ((long long)b * g.N + n) * C + c) : 0.f;
    }
    __syncthreads();
    for (int t = 0; t < XT; ++t) {
      const int n = t0 + t;
      if (n >= n1) break;
      const float* qh = qs + t * C + h0;
      const float* kh = ks + t * C + h0;
      const float* gh = gs + t * C + h0;
      float gq = -cqc * qs[t * C + c], gk = -ckc * ks[t * C + c], gv = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        gq = fmaf(mq[j], kh[j], gq);
        gk = fmaf(mqt[j], qh[j], gk);
        gv = fmaf(at[j], gh[j], gv);
      }
      float* orow = gqkv + ((long long)b * g.N + n) * 3 * C;
      orow[c] = gq, orow[C + c] = gk, orow[2 * C + c] = gv;
    }
  }
}


// ---- tensor-map TMA variants -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// same contract as xca_gram_kernel; a tile = box {C, XT} of map `ma` at column ax (and of `mb` at column bx), rows b * N + n
template <int D>
__global__ void xca_gram_tma_kernel(const __grid_constant__ CUtensorMap ma, int ax, const __grid_constant__ CUtensorMap mb, int bx, XcaTok g,
                                    float* __restrict__ partial) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bars[2];
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  const int tile = XT * g.C, ntiles = (n1 - n0 + XT - 1) / XT;
  const uint32_t bar0 = tc::smem_u32(bars), sm0 = tc::smem_u32(sm);
  if (c == 0) {
    tc::mbar_init(bar0, 1), tc::mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int i) {
    const uint32_t st = (uint32_t)(i & 1), bar = bar0 + 8 * st, dst = sm0 + st * 2u * tile * 4u;
    const int y = b * g.N + n0 + i * XT;
    tc::fence_async_smem();   // the stage's previous contents were read through the generic proxy
    mbar_expect_tx(bar, 2u * tile * 4u);
    tma_load_2d(dst, &ma, ax, y, bar);
    tma_load_2d(dst + tile * 4u, &mb, bx, y, bar);
  };
  if (c == 0 && ntiles > 0) issue(0);
  float G[D];
#pragma unroll
  for (int j = 0; j < D; ++j) G[j] = 0.f;
  float na = 0.f, nb = 0.f;
  for (int i = 0; i < ntiles; ++i) {
    if (c == 0 && i + 1 < ntiles) issue(i + 1);   // (its stage was released by the barrier that closed iteration i - 1)
    tc::mbar_wait(bar0 + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
    const float* as = sm + (i & 1) * 2 * tile;
    const float* bs = as + tile;
    const int tmax = min(XT, n1 - n0 - i * XT);   // rows past the chunk hold the next chunk's tokens (or zero fill): not summed
#pragma unroll 2
    for (int t = 0; t < tmax; ++t) {
      const float ai = as[t * g.C + c], bc = bs[t * g.C + c];
      na = fmaf(ai, ai, na), nb = fmaf(bc, bc, nb);
      const float* bh = bs + t * g.C + h0;
#pragma unroll
      for (int j = 0; j < D; ++j) G[j] = fmaf(ai, bh[j], G[j]);
    }
    __syncthreads();
  }
  float* p = partial + (((size_t)b * g.chunks + blockIdx.x) * g.C + c) * (D + 2);
#pragma unroll
  for (int j = 0; j < D; ++j) p[j] = G[j];
  p[D] = na, p[D + 1] = nb;
}

// same contract as xca_apply_kernel; source tiles = box {C, XT} of map `ms` at column sx
template <int D>
__global__ void xca_apply_tma_kernel(const float* __restrict__ M, const __grid_constant__ CUtensorMap ms, int sx, float* __restrict__ dst,
                                     long long dst_stride, XcaTok g) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bars[2];
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  const int tile = XT * g.C, ntiles = (n1 - n0 + XT - 1) / XT;
  const uint32_t bar0 = tc::smem_u32(bars), sm0 = tc::smem_u32(sm);
  if (c == 0) {
    tc::mbar_init(bar0, 1), tc::mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int i) {
    const uint32_t st = (uint32_t)(i & 1), bar = bar0 + 8 * st;
    tc::fence_async_smem();
    mbar_expect_tx(bar, (uint32_t)tile * 4u);
    tma_load_2d(sm0 + st * tile * 4u, &ms, sx, b * g.N + n0 + i * XT, bar);
  };
  if (c == 0 && ntiles > 0) issue(0);
  float m[D];
#pragma unroll
  for (int j = 0; j < D; ++j) m[j] = __ldg(M + ((size_t)b * g.C + c) * D + j);
  for (int i = 0; i < ntiles; ++i) {
    if (c == 0 && i + 1 < ntiles) issue(i + 1);
    tc::mbar_wait(bar0 + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
    const float* src = sm + (i & 1) * tile;
    const int t0 = n0 + i * XT, tmax = min(XT, n1 - t0);
#pragma unroll 2
    for (int t = 0; t < tmax; ++t) {
      const float* sh = src + t * g.C + h0;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(m[j], sh[j], s);
      dst[((long long)b * g.N + t0 + t) * dst_stride + c] = s;
    }
    __syncthreads();
  }
}

// TMA variant of xca_bwd_apply_kernel: q / k tiles from the qkv map (columns 0 and C), grad_out tiles from its own map
template <int D>
__global__ void xca_bwd_apply_tma_kernel(const __grid_constant__ CUtensorMap mqkv, const __grid_constant__ CUtensorMap mg, const float* __restrict__ Mq,
                                         const float* __restrict__ MqT, const float* __restrict__ AT, const float* __restrict__ cq,
                                         const float* __restrict__ ck, float* __restrict__ gqkv, XcaTok g) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bars[2];
  const int c = threadIdx.x, h0 = (c / D) * D, b = blockIdx.y, C = g.C;
  const int n0 = blockIdx.x * g.tokens_per_cta, n1 = min(g.N, n0 + g.tokens_per_cta);
  const int tile = XT * C, ntiles = (n1 - n0 + XT - 1) / XT;
  const uint32_t bar0 = tc::smem_u32(bars), sm0 = tc::smem_u32(sm);
  if (c == 0) {
    tc::mbar_init(bar0, 1), tc::mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int i) {
    const uint32_t st = (uint32_t)(i & 1), bar = bar0 + 8 * st, dst = sm0 + st * 3u * tile * 4u;
    const int y = b * g.N + n0 + i * XT;
    tc::fence_async_smem();
    mbar_expect_tx(bar, 3u * tile * 4u);
    tma_load_2d(dst, &mqkv, 0, y, bar);
    tma_load_2d(dst + tile * 4u, &mqkv, C, y, bar);
    tma_load_2d(dst + 2u * tile * 4u, &mg, 0, y, bar);
  };
  if (c == 0 && ntiles > 0) issue(0);
  float mq[D], mqt[D], at[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    mq[j] = __ldg(Mq + ((size_t)b * C + c) * D + j), mqt[j] = __ldg(MqT + ((size_t)b * C + c) * D + j);
    at[j] = __ldg(AT + ((size_t)b * C + c) * D + j);
  }
  const float cqc = __ldg(cq + (size_t)b * C + c), ckc = __ldg(ck + (size_t)b * C + c);
  for (int i = 0; i < ntiles; ++i) {
    if (c == 0 && i + 1 < ntiles) issue(i + 1);
    tc::mbar_wait(bar0 + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
    const float* qs = sm + (i & 1) * 3 * tile;
    const float* ks = qs + tile;
    const float* gs = ks + tile;
    const int t0 = n0 + i * XT, tmax = min(XT, n1 - t0);
    for (int t = 0; t < tmax; ++t) {
      const float* qh = qs + t * C + h0;
      const float* kh = ks + t * C + h0;
      const float* gh = gs + t * C + h0;
      float gq = -cqc * qs[t * C + c], gk = -ckc * ks[t * C + c], gv = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        gq = fmaf(mq[j], kh[j], gq);
        gk = fmaf(mqt[j], qh[j], gk);
        gv = fmaf(at[j], gh[j], gv);
      }
      float* orow = gqkv + ((long long)b * g.N + t0 + t) * 3 * C;
      orow[c] = gq, orow[C + c] = gk, orow[2 * C + c] = gv;
    }
    __syncthreads();
  }
}

// 2-D fp32 tensor map over a row-major matrix (rows x cols, row stride in floats), box = box_cols x XT rows, zero fill outside
static bool xca_make_map(CUtensorMap* map, const float* base, long long rows, int cols, long long row_stride, int box_cols) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("DD_NO_TMA") == nullptr && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    (void)cudaGetLastError();
  }
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)XT};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D>
static int xca_tma_configure() {
  static bool done = false;
  if (!done) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(xca_gram_tma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * XT * 256 * (int)sizeof(float)));
    DD_CHECK_CUDA(cudaFuncSetAttribute(xca_apply_tma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * XT * 256 * (int)sizeof(float)));
    DD_CHECK_CUDA(cudaFuncSetAttribute(xca_bwd_apply_tma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * XT * 256 * (int)sizeof(float)));
    done = true;
  }
  return DD_OK;
}
static int xca_tma_configure(int D) { return D == 8 ? xca_tma_configure<8>() : (D == 16 ? xca_tma_configure<16>() : xca_tma_configure<28>()); }

static int xca_tokens_per_cta(int B, int N) {
  // ~4 CTAs per SM over the batch, whole tiles, at most 256 tokens
  long long t = ((long long)B * N + 591) / 592;
  t = (t + XT - 1) / XT * XT;
  return (int)(t < XT ? XT : (t > 256 ? 256 : t));
}

struct XcaPlan {
  XcaTok g;
  size_t partial_floats, coef_floats;   // per-chunk partials; Mq | MqT | AT | cq | ck
};

static XcaPlan xca_plan(int B, int N, int C, int D) {
  XcaPlan p;
  p.g.N = N, p.g.C = C;
  p.g.tokens_per_cta = xca_tokens_per_cta(B, N);
  p.g.chunks = (N + p.g.tokens_per_cta - 1) / p.g.tokens_per_cta;
  p.partial_floats = (size_t)B * p.g.chunks * C * (D + 2);
  p.coef_floats = (size_t)B * C * (3 * D + 2);
  return p;
}

static int xca_check(const char* what, int B, int N, int C, int heads) {
  DD_REQUIRE(B > 0 && N > 0 && C > 0 && heads > 0 && C % heads == 0, "%s: bad shape B=%d N=%d C=%d heads=%d", what, B, N, C, heads);
  const int D = C / heads;
  DD_REQUIRE(D == 8 || D == 16 || D == 28, "%s: head dimension %d not built (8, 16, 28)", what, D);
  DD_REQUIRE(C % 32 == 0 && C <= 256, "%s: C must be a multiple of 32 and <= 256 (got %d)", what, C);
  DD_REQUIRE(B <= 65535, "%s: batch too large", what);
  return DD_OK;
}

#define DD_XCA_DISPATCH(D, KERNEL, ...)         \
  do {                                          \
    if (D == 8) KERNEL<8> __VA_ARGS__;          \
    else if (D == 16) KERNEL<16> __VA_ARGS__;   \
    else KERNEL<28> __VA_ARGS__;                \
  } while (0)

}  // namespace dd

extern "C" {

size_t dd_xca_workspace_bytes(int B, int N, int C, int heads) {
  if (B <= 0 || N <= 0 || C <= 0 || heads <= 0 || C % heads != 0) return 0;
  const dd::XcaPlan p = dd::xca_plan(B, N, C, C / heads);
  return (p.partial_floats + p.coef_floats) * sizeof(float);
}

int dd_xca_fwd(const float* qkv, const float* temperature, int B, int N, int C, int heads, float* out, float* attn, float* scores, float* rq,
               float* rk, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(qkv && temperature && out && attn && scores && rq && rk, "dd_xca_fwd: NULL pointer");
  if (int rc = xca_check("dd_xca_fwd", B, N, C, heads)) return rc;
  const int D = C / heads;
  const XcaPlan p = xca_plan(B, N, C, D);
  if (!workspace || workspace_bytes < p.partial_floats * sizeof(float)) {
    set_error("dd_xca_fwd: workspace too small (%zu < %zu)", workspace_bytes, p.partial_floats * sizeof(float));
    return DD_ERR_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(p.g.chunks, B);
  CUtensorMap mqkv;
  const bool tma = (((uintptr_t)qkv & 15) == 0) && xca_make_map(&mqkv, qkv, (long long)B * N, 3 * C, 3ll * C, C) && xca_tma_configure(D) == DD_OK;
  if (tma) DD_XCA_DISPATCH(D, xca_gram_tma_kernel, <<<grid, C, 4 * XT * C * sizeof(float), st>>>(mqkv, 0, mqkv, C, p.g, partial));
  else DD_XCA_DISPATCH(D, xca_gram_kernel, <<<grid, C, 2 * XT * C * sizeof(float), st>>>(qkv, 3ll * C, qkv + C, 3ll * C, p.g, partial));
  DD_XCA_DISPATCH(D, xca_softmax_kernel, <<<B, C, C * sizeof(float), st>>>(partial, p.g.chunks, C, temperature, attn, scores, rq, rk));
  if (tma) DD_XCA_DISPATCH(D, xca_apply_tma_kernel, <<<grid, C, 2 * XT * C * sizeof(float), st>>>(attn, mqkv, 2 * C, out, (long long)C, p.g));
  else DD_XCA_DISPATCH(D, xca_apply_kernel, <<<grid, C, XT * C * sizeof(float), st>>>(attn, qkv + 2 * C, 3ll * C, out, (long long)C, p.g));
  count_launches(3);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_xca_bwd(const float* qkv, const float* temperature, const float* grad_out, const float* attn, const float* scores, const float* rq,
               const float* rk, int B, int N, int C, int heads, float* grad_qkv, float* grad_temp_part, void* workspace, size_t workspace_bytes,
               void* stream) {
  using namespace dd;
  DD_REQUIRE(qkv && temperature && grad_out && attn && scores && rq && rk && grad_qkv && grad_temp_part, "dd_xca_bwd: NULL pointer");
  if (int rc = xca_check("dd_xca_bwd", B, N, C, heads)) return rc;
  const int D = C / heads;
  const XcaPlan p = xca_plan(B, N, C, D);
  if (!workspace || workspace_bytes < (p.partial_floats + p.coef_floats) * sizeof(float)) {
    set_error("dd_xca_bwd: workspace too small (%zu < %zu)", workspace_bytes, (p.partial_floats + p.coef_floats) * sizeof(float));
    return DD_ERR_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  float* Mq = partial + p.partial_floats;
  float* MqT = Mq + (size_t)B * C * D;
  float* AT = MqT + (size_t)B * C * D;
  float* cq = AT + (size_t)B * C * D;
  float* ck = cq + (size_t)B * C;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(p.g.chunks, B);
  CUtensorMap mqkv, mg;
  const bool tma = ((((uintptr_t)qkv | (uintptr_t)grad_out) & 15) == 0) && xca_make_map(&mqkv, qkv, (long long)B * N, 3 * C, 3ll * C, C) &&
                   xca_make_map(&mg, grad_out, (long long)B * N, C, (long long)C, C) && xca_tma_configure(D) == DD_OK;
  if (tma) DD_XCA_DISPATCH(D, xca_gram_tma_kernel, <<<grid, C, 4 * XT * C * sizeof(float), st>>>(mg, 0, mqkv, 2 * C, p.g, partial));
  else DD_XCA_DISPATCH(D, xca_gram_kernel, <<<grid, C, 2 * XT * C * sizeof(float), st>>>(grad_out, (long long)C, qkv + 2 * C, 3ll * C, p.g, partial));
  DD_XCA_DISPATCH(D, xca_softmax_bwd_kernel, <<<B, C, (C + C * (D + 1)) * sizeof(float), st>>>(partial, p.g.chunks, C, temperature, attn, scores, rq,
                                                                                             rk, Mq, MqT, AT, cq, ck, grad_temp_part));
  if (tma) DD_XCA_DISPATCH(D, xca_bwd_apply_tma_kernel, <<<grid, C, 6 * XT * C * sizeof(float), st>>>(mqkv, mg, Mq, MqT, AT, cq, ck, grad_qkv, p.g));
  else DD_XCA_DISPATCH(D, xca_bwd_apply_kernel, <<<grid, C, 3 * XT * C * sizeof(float), st>>>(qkv, grad_out, Mq, MqT, AT, cq, ck, grad_qkv, p.g));
  count_launches(3);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
