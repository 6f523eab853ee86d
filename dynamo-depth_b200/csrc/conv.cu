// Fused decoder convolutions: forward, data gradient, weight gradient (fp32 SIMT implicit GEMM).
//
//   out = act( conv_k( pad( concat( up(x0), x1 ) ) ) + bias ) [+ residual]
//
// The "virtual input" (nearest up-sampling, skip concatenation, reflection / zero padding) is never
// materialised: the tile loader evaluates it while staging shared memory (layers.py:85-121,
// depth_decoder.py:46-53,103-113, motion_decoder.py:34-62); only bilinear x2 up-sampling is written once
// to the workspace (materialise_up).  3x3 layers with more than 16 output channels run the Winograd
// F(2x2,3x3) kernel, 3x3 layers with at most 16 the small-Cout kernel, everything else the direct core:
// one CTA = 128 threads computes an 8x16 pixel tile for 8*CPT output channels, every thread owns CPT
// channels x 8 consecutive pixels.
// The data gradient reuses the same core on the (zero-extended) output gradient with flipped,
// transposed weights, producing the gradient on the padded grid; a light routing kernel then folds
// the reflection border back and transposes the up-sampling / concatenation.
#include <stdlib.h>
#include <string.h>

#include "dd_common.cuh"

namespace dd {

#ifndef DD_WINO_UNROLL
#define DD_WINO_UNROLL 4   // K-loop unrolling of the Winograd GEMMs (operand double-buffering vs register pressure)
#endif
constexpr int WINO_UNROLL = DD_WINO_UNROLL;
constexpr int CT_H = 8, CT_W = 16;
constexpr int CI_T = 8;
constexpr int IN_PITCH = 20;
constexpr int IN_PLANE = (CT_H + 2) * IN_PITCH + 4;   // 204: staggers channel planes across banks
constexpr int CONV_THREADS = 128;

struct VirtIn {
  const float* x0;
  const float* x1;
  int C0, C1, H0, W0, up0;
  int Hin, Win;
  int pad_mode;
};

// Source of one position of an input tile, independent of the channel: derived once per thread (registers) so that
// the per-K-step staging is an address add + load instead of re-deriving padding, reflection and up-sampling taps for
// every element of every channel.  (Bilinear x2 up-sampling is materialised by the host wrapper before the conv
// kernels run, so only o00 is live there; the four-tap form remains for completeness of the virtual-input model.)
struct TapEntry {
  int o00, o01, o10, o11;   // offsets inside one x0 channel plane (o00 < 0: zero padding / outside)
  float ly, lx;             // bilinear weights (DD_UP_BILINEAR2 only)
};

__device__ __forceinline__ void build_tile_map(const VirtIn& v, int y, int x, TapEntry& e0, int& o1) {
  e0.o00 = e0.o01 = e0.o10 = e0.o11 = -1;
  e0.ly = e0.lx = 0.f;
  o1 = -1;
  if (y < -1 || y > v.Hin || x < -1 || x > v.Win) return;
  if (v.pad_mode == DD_PAD_REFLECT) {
    y = reflect1(y, v.Hin);
    x = reflect1(x, v.Win);
  } else if (y < 0 || y >= v.Hin || x < 0 || x >= v.Win) {
    return;
  }
  o1 = y * v.Win + x;
  if (v.up0 == DD_UP_NONE) {
    e0.o00 = y * v.W0 + x;
  } else if (v.up0 == DD_UP_NEAREST2) {
    e0.o00 = (y >> 1) * v.W0 + (x >> 1);
  } else {
    const Taps ty = up_taps(y, 1, v.H0), tx = up_taps(x, 1, v.W0);
    e0.o00 = ty.i0 * v.W0 + tx.i0, e0.o01 = ty.i0 * v.W0 + tx.i1;
    e0.o10 = ty.i1 * v.W0 + tx.i0, e0.o11 = ty.i1 * v.W0 + tx.i1;
    e0.ly = ty.l, e0.lx = tx.l;
  }
}

struct ConvArgs {
  VirtIn vin;
  int B, Ho, Wo;    // output grid
  int oy, ox;       // virtual-input coordinate of the centre tap of output (0,0)
  int Cin, Cout;
  const float* wt;  // prepared weights [Cin][k*k][cout_pad]
  int cout_pad;
  const float* bias;
  const float* residual;
  int act;
  float* out;
  float* out1;      // channels >= split go to out1 (data gradient written straight into grad_x0 / grad_x1)
  int split;        // 0 = single output tensor
  int tiles_x;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DD_ACT_ELU) return v > 0.f ? v : expf(v) - 1.f;   // nn.ELU(alpha=1) (layers.py:92)
  if (act == DD_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == DD_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

template <int KS, int CPT>
__global__ void __launch_bounds__(CONV_THREADS, 4) conv_core_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int KK = KS * KS;
  constexpr int HALO = KS / 2;
  constexpr int ROWS = CT_H + 2 * HALO, COLS = CT_W + 2 * HALO;
  constexpr int NPOS = ROWS * COLS;
  constexpr int CO_T = 8 * CPT;
  constexpr int NV = 8 + KS - 1;
  constexpr int NSLOT = (NPOS + CONV_THREADS - 1) / CONV_THREADS;   // input-tile positions per thread (2 for 3x3, 1 for 1x1)
  __shared__ __align__(16) float in_s[2][CI_T * IN_PLANE];
  __shared__ __align__(16) float w_s[CI_T * KK * CO_T];

  const int tid = threadIdx.x;
  const int pg = tid & 15, cg = tid >> 4;
  const int row = pg & 7, seg = pg >> 3;
  const int tile = blockIdx.x;
  const int ty0 = (tile / a.tiles_x) * CT_H, tx0 = (tile % a.tiles_x) * CT_W;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;

  // Loader role: thread = fixed position(s) of the input tile, all 8 channels of a K-step.  The position's source
  // (padding, reflection, nearest up-sampling) is channel and step independent: derived once, kept in registers as
  // running pointers that advance one plane per channel (bilinear up-sampling is materialised by the host wrapper).
  const int C0 = a.vin.C0, Cin = a.Cin;
  const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
  const float* src0[NSLOT];
  const float* src1[NSLOT];
  bool ok0[NSLOT], ok1[NSLOT];
  int s_off[NSLOT];
#pragma unroll
  for (int sl = 0; sl < NSLOT; ++sl) {
    const int pos = tid + sl * CONV_THREADS;
    const int r = pos / COLS, c = pos - r * COLS;
    TapEntry te;
    int o1;
    build_tile_map(a.vin, ty0 + r - HALO + a.oy, tx0 + c - HALO + a.ox, te, o1);
    ok0[sl] = pos < NPOS && te.o00 >= 0, ok1[sl] = pos < NPOS && o1 >= 0 && a.vin.x1 != nullptr;
    s_off[sl] = pos < NPOS ? r * IN_PITCH + c : -1;
    src0[sl] = a.vin.x0 + (size_t)b * C0 * plane0 + (ok0[sl] ? te.o00 : 0);
    src1[sl] = ok1[sl] ? a.vin.x1 + ((ptrdiff_t)b * a.vin.C1 - C0) * (ptrdiff_t)plane1 + o1 : a.vin.x0;
  }

  // accumulators: channel pairs packed for FFMA2 (CPT >= 2), scalar for the single-channel variant
  constexpr int CP2 = CPT >= 2 ? CPT / 2 : 1;
  f32x2 acc2[CP2][8];
  float acc1[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    acc1[p] = 0.f;
#pragma unroll
    for (int c = 0; c < CP2; ++c) acc2[c][p] = 0ull;
  }

  // K-steps are visited in order: every call advances the running pointers by 8 planes
  auto gather = [&](int ci0, float (&pre)[NSLOT][CI_T]) {
#pragma unroll
    for (int ci = 0; ci < CI_T; ++ci) {
      const int c = ci0 + ci;   // CTA-uniform
      const bool in0 = c < C0;
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const float* ptr = in0 ? src0[sl] : src1[sl];
        const bool ok = in0 ? ok0[sl] : (ok1[sl] && c < Cin);
        float v = 0.f;
        if (ok) v = __ldg(ptr);
        pre[sl][ci] = v;
        src0[sl] += plane0, src1[sl] += plane1;
      }
    }
  };
  auto scatter = [&](int buf, const float (&pre)[NSLOT][CI_T]) {
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl)
      if (s_off[sl] >= 0) {
#pragma unroll
        for (int ci = 0; ci < CI_T; ++ci) in_s[buf][ci * IN_PLANE + s_off[sl]] = pre[sl][ci];
      }
  };
  auto stage_weights = [&](int ci0) {
    constexpr int PER = CO_T / 4;
    for (int i = tid; i < CI_T * KK * PER; i += CONV_THREADS) {
      const int q = i % PER, ct = i / PER;   // ct = ci*KK + tap
      const int ci = ct / KK;
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ci0 + ci < a.Cin)
        w4 = __ldg(reinterpret_cast<const float4*>(a.wt + ((size_t)(ci0 * KK + ct)) * a.cout_pad + co0 + q * 4));
      reinterpret_cast<float4*>(w_s)[ct * PER + q] = w4;
    }
  };

  float pre[NSLOT][CI_T];
  gather(0, pre);
  scatter(0, pre);

  const int nk = (a.Cin + CI_T - 1) / CI_T;
  for (int k = 0; k < nk; ++k) {
    const int buf = k & 1;
    stage_weights(k * CI_T);
    __syncthreads();   // in_s[buf] and w_s of step k are visible
    if (k + 1 < nk) gather((k + 1) * CI_T, pre);   // loads in flight during the FMAs below
    const float* in_b = in_s[buf];
#pragma unroll 1
    for (int ci = 0; ci < CI_T; ++ci) {
      const float* ip = in_b + ci * IN_PLANE + row * IN_PITCH + seg * 8;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        float iv[NV];
        const float4 v0 = *reinterpret_cast<const float4*>(ip + dy * IN_PITCH);
        const float4 v1 = *reinterpret_cast<const float4*>(ip + dy * IN_PITCH + 4);
        iv[0] = v0.x, iv[1] = v0.y, iv[2] = v0.z, iv[3] = v0.w;
        iv[4] = v1.x, iv[5] = v1.y, iv[6] = v1.z, iv[7] = v1.w;
        if (KS == 3) {
          const float2 v2 = *reinterpret_cast<const float2*>(ip + dy * IN_PITCH + 8);
          iv[NV - 2] = v2.x, iv[NV - 1] = v2.y;
        }
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float* wp = w_s + (ci * KK + dy * KS + dx) * CO_T + cg * CPT;
          if (CPT >= 2) {   // (w[c], w[c+1]) * broadcast(iv) + (acc[c], acc[c+1]): one FFMA2 per channel pair and pixel
            f32x2 wv2[CP2];
            if (CPT >= 4) {
#pragma unroll
              for (int q = 0; q < CPT / 4; ++q) {
                const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wp + q * 4);
                wv2[q * 2 + 0] = w4.x, wv2[q * 2 + 1] = w4.y;
              }
            } else {
              wv2[0] = *reinterpret_cast<const f32x2*>(wp);
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const f32x2 ivv = pack2(iv[p + dx], iv[p + dx]);
#pragma unroll
              for (int c = 0; c < CP2; ++c) acc2[c][p] = fma2(wv2[c], ivv, acc2[c][p]);
            }
          } else {
            const float wv = wp[0];
#pragma unroll
            for (int p = 0; p < 8; ++p) acc1[p] = fmaf(wv, iv[p + dx], acc1[p]);
          }
        }
      }
    }
    if (k + 1 < nk) scatter(buf ^ 1, pre);   // other buffer: nobody reads it during this step
    __syncthreads();                         // all reads of w_s / in_s[buf] done before they are overwritten
  }

  // epilogue: bias, activation, residual, store
  const int y = ty0 + row;
  const int xb = tx0 + seg * 8;
  if (y >= a.Ho) return;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int co = co0 + cg * CPT + c;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    float* outp = a.out;
    size_t o = (((size_t)b * a.Cout + co) * a.Ho + y) * a.Wo + xb;
    if (a.split > 0) {   // two destination tensors with split / Cout-split channels
      if (co < a.split) o = (((size_t)b * a.split + co) * a.Ho + y) * a.Wo + xb;
      else outp = a.out1, o = (((size_t)b * (a.Cout - a.split) + (co - a.split)) * a.Ho + y) * a.Wo + xb;
      if (outp == nullptr) continue;
    }
    float v[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float accv;
      if (CPT >= 2) {
        float lo, hi;
        unpack2(acc2[c >> 1][p], lo, hi);
        accv = (c & 1) ? hi : lo;
      } else {
        accv = acc1[p];
      }
      v[p] = apply_act(accv + bv, a.act);
      if (a.residual && xb + p < a.Wo) v[p] += __ldg(a.residual + o + p);
    }
    if (xb + 7 < a.Wo && (a.Wo & 3) == 0) {
      reinterpret_cast<float4*>(outp + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(outp + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (xb + p < a.Wo) outp[o + p] = v[p];
    }
  }
}

// ---- 3x3 layers with few output channels (disparity heads, full-resolution motion level, their data gradients) ----
// The 8x16-pixel x 8-channel-group mapping of conv_core_kernel leaves most lanes idle when Cout <= 16, and these layers
// are bound by the input stream, not by FMAs.  Here every thread owns 4 consecutive pixels x ALL output channels
// (CO = 2/4/12/16, FFMA2 on channel pairs), reads its 3x6 input window straight from global memory (neighbouring
// threads share the lines through L1) and takes the weights of a 16-channel chunk from shared memory.
// CTA = 256 threads = 16 rows x 64 columns of the output.  Inputs without up-sampling only (x0 and x1 share the grid).
constexpr int SM_TH = 16, SM_TW = 64, SM_THREADS = 256, SM_CI = 16;

template <int CO>
__global__ void __launch_bounds__(SM_THREADS, 2) conv_small_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int CP = CO / 2;
  __shared__ __align__(16) float w_s[SM_CI * 9 * CO];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = (tid & 15) * 4;
  const int tile = blockIdx.x;
  const int y = (tile / a.tiles_x) * SM_TH + ty, x0p = (tile % a.tiles_x) * SM_TW + tx;
  const int b = blockIdx.z;
  const int Hin = a.vin.Hin, Win = a.vin.Win, C0 = a.vin.C0, Cin = a.Cin;
  const size_t plane = (size_t)Hin * Win;
  const bool reflect = a.vin.pad_mode == DD_PAD_REFLECT;

  // source offsets of the 3 rows / 6 columns this thread reads (channel independent); -1 = zero (padding / overhang)
  int ro[3], co[6];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int yy = y + dy - 1 + a.oy;
    if (reflect && yy >= -1 && yy <= Hin) yy = reflect1(yy, Hin);
    ro[dy] = (yy >= 0 && yy < Hin) ? yy * Win : -1;
  }
#pragma unroll
  for (int dx = 0; dx < 6; ++dx) {
    int xx = x0p + dx - 1 + a.ox;
    if (reflect && xx >= -1 && xx <= Win) xx = reflect1(xx, Win);
    co[dx] = (xx >= 0 && xx < Win) ? xx : -1;
  }

  f32x2 acc[CP][4];
#pragma unroll
  for (int c = 0; c < CP; ++c)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[c][p] = 0ull;

  const float* src0 = a.vin.x0 + (size_t)b * C0 * plane;
  const float* src1 = a.vin.x1 ? a.vin.x1 + ((ptrdiff_t)b * a.vin.C1 - C0) * (ptrdiff_t)plane : a.vin.x0;
  for (int ci0 = 0; ci0 < Cin; ci0 += SM_CI) {
    __syncthreads();   // previous chunk's weights no longer read
    const int chunk_floats = min(SM_CI, Cin - ci0) * 9 * CO;   // weights of channels >= Cin are never read below
    for (int i = tid; i < SM_CI * 9 * CO / 4; i += SM_THREADS) {
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i * 4 < chunk_floats) w4 = __ldg(reinterpret_cast<const float4*>(a.wt + (size_t)ci0 * 9 * CO) + i);
      reinterpret_cast<float4*>(w_s)[i] = w4;
    }
    __syncthreads();
    const int nci = min(SM_CI, Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const int c = ci0 + ci;
      const float* pl = (c < C0 ? src0 : src1) + (size_t)c * plane;
      const float* wc = w_s + ci * 9 * CO;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        float v[6];
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) v[dx] = (ro[dy] >= 0 && co[dx] >= 0) ? __ldg(pl + ro[dy] + co[dx]) : 0.f;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          f32x2 w2[CP];
#pragma unroll
          for (int q = 0; q < CP; ++q) w2[q] = *reinterpret_cast<const f32x2*>(wc + (dy * 3 + dx) * CO + 2 * q);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const f32x2 vv = pack2(v[p + dx], v[p + dx]);
#pragma unroll
            for (int q = 0; q < CP; ++q) acc[q][p] = fma2(w2[q], vv, acc[q][p]);
          }
        }
      }
    }
  }

  if (y >= a.Ho || x0p >= a.Wo) return;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    if (c >= a.Cout) continue;
    const float bv = a.bias ? __ldg(a.bias + c) : 0.f;
    float* outp = a.out;
    size_t o = (((size_t)b * a.Cout + c) * a.Ho + y) * a.Wo + x0p;
    if (a.split > 0) {   // two destination tensors with split / Cout-split channels
      if (c < a.split) o = (((size_t)b * a.split + c) * a.Ho + y) * a.Wo + x0p;
      else outp = a.out1, o = (((size_t)b * (a.Cout - a.split) + (c - a.split)) * a.Ho + y) * a.Wo + x0p;
      if (outp == nullptr) continue;
    }
    float r[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float lo, hi;
      unpack2(acc[c >> 1][p], lo, hi);
      r[p] = apply_act(((c & 1) ? hi : lo) + bv, a.act);
      if (a.residual && x0p + p < a.Wo) r[p] += __ldg(a.residual + o + p);
    }
    if (x0p + 3 < a.Wo && (a.Wo & 3) == 0) {
      *reinterpret_cast<float4*>(outp + o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (x0p + p < a.Wo) outp[o + p] = r[p];
    }
  }
}

// wt[ci][tap][co] (zero padded to cout_pad) from OIHW weights; `transpose` builds the data-gradient
// operator: wt[co_f][8-tap][ci_f] (flipped taps, swapped channel roles).
__global__ void conv_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int KK,
                                         int out_pad, int transpose) {
  const int n_in = transpose ? Cout : Cin;    // rows of wt
  const int n_out = transpose ? Cin : Cout;   // valid columns of wt
  const size_t total = (size_t)n_in * KK * out_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % out_pad);
    const int tap = (int)((i / out_pad) % KK);
    const int rowi = (int)(i / ((size_t)out_pad * KK));
    float v = 0.f;
    if (col < n_out) {
      if (!transpose) v = __ldg(w + ((size_t)col * Cin + rowi) * KK + tap);
      else v = __ldg(w + ((size_t)rowi * Cin + col) * KK + (KK - 1 - tap));
    }
    wt[i] = v;
  }
}

// g_conv = grad_out * act'(out)
__global__ void conv_act_grad_kernel(const float* __restrict__ grad_out, const float* __restrict__ out, float* __restrict__ g,
                                     size_t n, int act) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float o = __ldg(out + i), go = __ldg(grad_out + i);
    float d = 1.f;
    if (act == DD_ACT_ELU) d = o > 0.f ? 1.f : o + 1.f;
    else if (act == DD_ACT_SIGMOID) d = o * (1.f - o);
    else if (act == DD_ACT_RELU) d = o > 0.f ? 1.f : 0.f;
    g[i] = go * d;
  }
}

struct RouteArgs {
  const float* gpad;   // (B, Cin, Hp, Wp) gradient on the (padded) convolution-input grid
  int B, Cin, H, W;    // virtual input size
  int Hp, Wp, off;     // padded grid and offset of virtual (0,0) inside it (1 reflect, 0 zero)
  int reflect;
  int C0, C1, H0, W0, up0;
  float* gx0;
  float* gx1;
};

__device__ __forceinline__ float folded(const RouteArgs& a, const float* __restrict__ plane, int y, int x) {
  // sum over the padded positions that ReflectionPad2d(1) maps onto (y, x): itself, -1 when the
  // coordinate is 1, size when it is size-2 (both for size == 3)
  int ys[3], xs[3];
  int ny = 0, nx = 0;
  ys[ny++] = y;
  xs[nx++] = x;
  if (a.reflect) {
    if (y == 1) ys[ny++] = -1;
    if (y == a.H - 2) ys[ny++] = a.H;
    if (x == 1) xs[nx++] = -1;
    if (x == a.W - 2) xs[nx++] = a.W;
  }
  float s = 0.f;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) s += __ldg(plane + (size_t)(ys[i] + a.off) * a.Wp + xs[j] + a.off);
  return s;
}

__global__ void conv_route_kernel(const __grid_constant__ RouteArgs a) {
  const size_t n0 = a.gx0 ? (size_t)a.B * a.C0 * a.H0 * a.W0 : 0;
  const size_t n1 = a.gx1 ? (size_t)a.B * a.C1 * a.H * a.W : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n0) {
      const int x = (int)(i % a.W0), y = (int)((i / a.W0) % a.H0);
      const int c = (int)((i / ((size_t)a.W0 * a.H0)) % a.C0), b = (int)(i / ((size_t)a.W0 * a.H0 * a.C0));
      const float* plane = a.gpad + ((size_t)b * a.Cin + c) * a.Hp * a.Wp;
      float s = 0.f;
      if (a.up0 == DD_UP_NONE) {
        s = folded(a, plane, y, x);
      } else if (a.up0 == DD_UP_NEAREST2) {
        s = folded(a, plane, 2 * y, 2 * x) + folded(a, plane, 2 * y, 2 * x + 1) + folded(a, plane, 2 * y + 1, 2 * x) +
            folded(a, plane, 2 * y + 1, 2 * x + 1);
      } else if (2 * y - 1 >= 2 && 2 * y + 2 <= a.H - 3 && 2 * x - 1 >= 2 && 2 * x + 2 <= a.W - 3) {
        // interior: none of the sixteen up-sampled positions that read this texel is a border, a reflection-fold row / column
        // or clamped, so the transposed bilinear x2 weights are the constants (1/4, 3/4, 3/4, 1/4) per axis
        const float* p0 = plane + (size_t)(2 * y - 1 + a.off) * a.Wp + (2 * x - 1 + a.off);
        const float wgt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
          const float* pr = p0 + (size_t)dy * a.Wp;
          const float row = 0.25f * __ldg(pr) + 0.75f * __ldg(pr + 1) + 0.75f * __ldg(pr + 2) + 0.25f * __ldg(pr + 3);
          s = fmaf(wgt[dy], row, s);
        }
      } else {
        for (int yy = 2 * y - 1; yy <= 2 * y + 2; ++yy) {
          if (yy < 0 || yy >= a.H) continue;
          const float wy = up_weight(yy, 1, a.H0, y);
          if (wy == 0.f) continue;
          for (int xx = 2 * x - 1; xx <= 2 * x + 2; ++xx) {
            if (xx < 0 || xx >= a.W) continue;
            const float wx = up_weight(xx, 1, a.W0, x);
            if (wx != 0.f) s += wy * wx * folded(a, plane, yy, xx);
          }
        }
      }
      a.gx0[i] = s;
    } else {
      const size_t j = i - n0;
      const int x = (int)(j % a.W), y = (int)((j / a.W) % a.H);
      const int c = (int)((j / ((size_t)a.W * a.H)) % a.C1), b = (int)(j / ((size_t)a.W * a.H * a.C1));
      const float* plane = a.gpad + ((size_t)b * a.Cin + a.C0 + c) * a.Hp * a.Wp;
      a.gx1[j] = folded(a, plane, y, x);
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int nbytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem_src), "r"(nbytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// ---- Winograd F(2x2, 3x3) core (fp32, CUDA cores) -------------------------------------------------
// 2.25x fewer multiply-adds than the direct form for the 3x3 layers: per CTA 8x16 output pixels = 32 tiles of
// 2x2, 32 output channels, K-steps of 8 input channels.  Each step transforms the staged input tile
// (V = B^T d B, 16 frequency planes) and runs 16 independent [32 co x 8 ci] x [8 ci x 32 tiles] products,
// thread = (frequency, 8 channels, 8 tiles); the epilogue applies Y = A^T m A through shared memory.
// Weights arrive pre-transformed (U = G g G^T) as wt[16][Cin][cout_pad].
constexpr int WN_THREADS = 256;
constexpr int WN_CO = 32;
constexpr int WN_TILES = 32;
constexpr int WN_VP = 36;    // V_s pitch over tiles
constexpr int WN_MP = 32;    // M_s pitch over tiles (columns XOR-swizzled by the channel group: conflict-free 64-bit stores)
constexpr int WN_IN_PITCH = 24;   // rows 2*ttr of the 4 tile rows land in disjoint 128-byte halves: conflict-free 64-bit transform loads
constexpr int WN_IN_PLANE = (CT_H + 2) * WN_IN_PITCH + 4;   // 244
constexpr int WN_SMEM_LOOP = 2 * CI_T * WN_IN_PLANE + 2 * 16 * CI_T * WN_VP + 2 * 16 * CI_T * WN_CO;
constexpr int WN_SMEM_EPI = 16 * WN_CO * WN_MP;
constexpr int WN_SMEM_FLOATS = WN_SMEM_LOOP > WN_SMEM_EPI ? WN_SMEM_LOOP : WN_SMEM_EPI;

__device__ __forceinline__ void emit_output(const ConvArgs& a, int b, int co, int y, int x, float v) {
  if (y >= a.Ho || x >= a.Wo || co >= a.Cout) return;
  v = apply_act(v + (a.bias ? __ldg(a.bias + co) : 0.f), a.act);
  float* outp = a.out;
  size_t o = (((size_t)b * a.Cout + co) * a.Ho + y) * a.Wo + x;
  if (a.split > 0) {
    if (co < a.split) o = (((size_t)b * a.split + co) * a.Ho + y) * a.Wo + x;
    else outp = a.out1, o = (((size_t)b * (a.Cout - a.split) + (co - a.split)) * a.Ho + y) * a.Wo + x;
    if (outp == nullptr) return;
  }
  if (a.residual) v += __ldg(a.residual + o);
  outp[o] = v;
}

}  // namespace dd

#include "conv_tc4.cuh"  // tcgen05 implicit-GEMM core of the 3x3 layers (uses ConvArgs, build_tile_map, emit_output from above)

namespace dd {

// x0 arrives either as is or nearest-up-sampled (one tap per element); bilinear up-sampling is materialised by the host
// wrapper (materialise_up) before the launch.
__global__ void __launch_bounds__(WN_THREADS, 2) conv_wino_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int ROWS = CT_H + 2, COLS = CT_W + 2, NPOS = ROWS * COLS;
  extern __shared__ __align__(16) float wsm[];
  float* in_s = wsm;                                  // [2][CI_T*WN_IN_PLANE]
  float* V_s = wsm + 2 * CI_T * WN_IN_PLANE;             // [2][16][CI_T][WN_VP], transformed one step ahead
  float* U_s = V_s + 2 * 16 * CI_T * WN_VP;           // [2][16][CI_T][WN_CO], filled by cp.async one step ahead
  float* M_s = wsm;                                   // epilogue only: [16][WN_CO][WN_MP]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int ty0 = (tile / a.tiles_x) * CT_H, tx0 = (tile % a.tiles_x) * CT_W;
  const int co0 = blockIdx.y * WN_CO;
  const int b = blockIdx.z;
  // GEMM role: frequency p, 8 output channels (cg), 8 tiles (tg)
  const int p = 2 * warp + (lane >> 4), cg = (lane >> 2) & 3, tg = lane & 3;
  // transform role: input channel ti, winograd tile tt
  const int ti = tid >> 5, tt = tid & 31, ttr = tt >> 3, ttc = tt & 7;

  // Loader role: thread = one position of the 10x18 input tile, all 8 channels of a K-step.  Where that position
  // reads from (padding, reflection, up-sampling taps) does not depend on the channel or the step, so it is derived
  // once and kept in registers; per element the loader is then an address add + load.
  const bool has_pos = tid < NPOS;
  TapEntry te;
  int o1;
  const int lr = tid / COLS, lc = tid - lr * COLS;
  build_tile_map(a.vin, ty0 + lr - 1 + a.oy, tx0 + lc - 1 + a.ox, te, o1);
  if (!has_pos) te.o00 = -1, o1 = -1;
  const int s_off = has_pos ? lr * WN_IN_PITCH + lc : 0;
  const int C0 = a.vin.C0, Cin = a.Cin;
  const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
  const bool ok0 = te.o00 >= 0, ok1 = o1 >= 0;
  // running sources of this thread's position: channel c of the concatenated input lives at src0 + c*plane0 (c < C0)
  // or src1 + c*plane1 (c >= C0; src1 is pre-shifted by -C0 planes); both advance by 8 planes per K-step
  const float* src0 = a.vin.x0 + (size_t)b * C0 * plane0 + (ok0 ? te.o00 : 0);
  const float* src1 = a.vin.x1 ? a.vin.x1 + ((ptrdiff_t)b * a.vin.C1 - C0) * (ptrdiff_t)plane1 + (ok1 ? o1 : 0) : a.vin.x0;

  f32x2 acc2[8][4];   // [channel][tile pair]: packed accumulators (FFMA2)
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc2[c][t] = 0ull;

  // gathers the 8 channels of the K-step starting at channel ci0 (steps are visited in order: the running pointers
  // advance by one plane per channel)
  auto gather = [&](int ci0, float (&pre)[CI_T]) {
#pragma unroll
    for (int ci = 0; ci < CI_T; ++ci) {
      const int c = ci0 + ci;   // CTA-uniform
      const bool in0 = c < C0;
      const float* ptr = in0 ? src0 : src1;
      const bool ok = in0 ? ok0 : (ok1 && c < Cin);
      float v = 0.f;
      if (ok) v = __ldg(ptr);
      pre[ci] = v;
      src0 += plane0, src1 += plane1;
    }
  };
  auto scatter = [&](int buf, const float (&pre)[CI_T]) {
    if (has_pos) {
      float* dst = in_s + buf * CI_T * WN_IN_PLANE + s_off;
#pragma unroll
      for (int ci = 0; ci < CI_T; ++ci) dst[ci * WN_IN_PLANE] = pre[ci];
    }
  };

  // U_s[buf][p][ci][co] <- wt[p][ci0+ci][co0 + co]: 16-byte chunk i = tid + 256*j -> (p = tid/64 + 4j, ci = tid/8 % 8, q = tid % 8)
  const int w_ci = (tid >> 3) & 7;
  const float* w_src = a.wt + ((size_t)(tid >> 6) * Cin + w_ci) * a.cout_pad + co0 + (tid & 7) * 4;
  const size_t w_jstride = (size_t)4 * Cin * a.cout_pad;
  auto issue_weights = [&](int ci0, int buf) {
    const bool ok = ci0 + w_ci < Cin;
    const float* src = ok ? w_src + (size_t)ci0 * a.cout_pad : a.wt;
    float* dst = U_s + buf * 16 * CI_T * WN_CO + tid * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) cp_async16(dst + j * 4 * WN_THREADS, ok ? src + j * w_jstride : src, ok);
    cp_async_commit();
  };

  // input transform V = B^T d B of (channel ti, tile tt): in_s[ib] -> V_s[vb]
  auto transform = [&](int ib, int vb) {
    const float* dp = in_s + ib * CI_T * WN_IN_PLANE + ti * WN_IN_PLANE + (2 * ttr) * WN_IN_PITCH + 2 * ttc;
    float d[4][4], t[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 lo = *reinterpret_cast<const float2*>(dp + i * WN_IN_PITCH);
      const float2 hi = *reinterpret_cast<const float2*>(dp + i * WN_IN_PITCH + 2);
      d[i][0] = lo.x, d[i][1] = lo.y, d[i][2] = hi.x, d[i][3] = hi.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[0][j] = d[0][j] - d[2][j];
      t[1][j] = d[1][j] + d[2][j];
      t[2][j] = d[2][j] - d[1][j];
      t[3][j] = d[1][j] - d[3][j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* vp = V_s + vb * 16 * CI_T * WN_VP + ((i * 4) * CI_T + ti) * WN_VP + tt;
      vp[0 * CI_T * WN_VP] = t[i][0] - t[i][2];
      vp[1 * CI_T * WN_VP] = t[i][1] + t[i][2];
      vp[2 * CI_T * WN_VP] = t[i][2] - t[i][1];
      vp[3 * CI_T * WN_VP] = t[i][1] - t[i][3];
    }
  };
  // 16 independent [32 co x 8 ci] x [8 ci x 32 tiles] products of one K-step; thread = (frequency, 8 channels, 8 tiles)
  auto gemm = [&](int buf) {
    const float* up = U_s + buf * 16 * CI_T * WN_CO + p * CI_T * WN_CO + cg * 8;
    const float* vp = V_s + buf * 16 * CI_T * WN_VP + p * CI_T * WN_VP + tg * 8;
#pragma unroll WINO_UNROLL
    for (int ci = 0; ci < CI_T; ++ci) {
      const float4 u0 = *reinterpret_cast<const float4*>(up + ci * WN_CO), u1 = *reinterpret_cast<const float4*>(up + ci * WN_CO + 4);
      const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(vp + ci * WN_VP), v1 = *reinterpret_cast<const ulonglong2*>(vp + ci * WN_VP + 4);
      const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      const f32x2 v[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const f32x2 uu = pack2(u[c], u[c]);
#pragma unroll
        for (int tq = 0; tq < 4; ++tq) acc2[c][tq] = fma2(uu, v[tq], acc2[c][tq]);
      }
    }
  };

  // Software pipeline over K-steps of 8 input channels, ONE barrier per step:
  //   step k:  cp.async U(k+1) | LDG tile(k+2) -> registers | transform(k+1): in_s -> V_s | GEMM(k) | registers -> in_s(k+2)
  // The transform and the GEMM of one step are independent; half of the warps run them in the opposite order so that
  // every SM sub-partition always has a warp on the FMA pipe while another one does the (FMA-free) transform.
  const int nk = (a.Cin + CI_T - 1) / CI_T;
  const bool gemm_first = (warp & 4) != 0;
  float pre[CI_T];
  issue_weights(0, 0);
  gather(0, pre);
  scatter(0, pre);
  __syncthreads();   // in_s[0] visible
  if (nk > 1) gather(CI_T, pre);
  transform(0, 0);
  if (nk > 1) scatter(1, pre);
  cp_async_wait_all();
  __syncthreads();   // V_s[0], U_s[0], in_s[1] visible

  for (int k = 0; k < nk; ++k) {
    const int buf = k & 1;
    const bool next = k + 1 < nk, next2 = k + 2 < nk;
    if (next) issue_weights((k + 1) * CI_T, buf ^ 1);
    if (next2) gather((k + 2) * CI_T, pre);   // loads in flight during the arithmetic below
    if (gemm_first) {
      gemm(buf);
      if (next) transform(buf ^ 1, buf ^ 1);
    } else {
      if (next) transform(buf ^ 1, buf ^ 1);
      gemm(buf);
    }
    if (next2) scatter(buf, pre);   // in_s[buf] was last read by transform(k) during step k-1
    cp_async_wait_all();
    __syncthreads();   // V_s[buf^1], U_s[buf^1], in_s[buf] complete; reads of V_s[buf] / U_s[buf] / in_s[buf^1] done
  }

  // epilogue: gather the 16 frequencies of every (channel, tile) through shared memory, Y = A^T m A
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int tq = 0; tq < 4; ++tq) {
      float lo, hi;
      unpack2(acc2[c][tq], lo, hi);
      *reinterpret_cast<float2*>(M_s + (p * WN_CO + cg * 8 + c) * WN_MP + ((tg * 8 + 2 * tq) ^ (2 * cg))) = make_float2(lo, hi);
    }
  __syncthreads();
  for (int q = tid; q < WN_CO * WN_TILES; q += WN_THREADS) {
    const int col = q >> 5, tl = q & 31;
    const int tr = tl >> 3, tc = tl & 7;
    const int y = ty0 + 2 * tr, x = tx0 + 2 * tc, co = co0 + col;
    if (co >= a.Cout || y >= a.Ho || x >= a.Wo) continue;
    float m[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) m[i][j] = M_s[((i * 4 + j) * WN_CO + col) * WN_MP + (tl ^ (2 * ((col >> 3) & 3)))];
    float sr[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sr[0][j] = m[0][j] + m[1][j] + m[2][j];
      sr[1][j] = m[1][j] - m[2][j] - m[3][j];
    }
    // destination of this channel (single tensor, or the two tensors of a directly written data gradient)
    float* outp = a.out;
    size_t plane = ((size_t)b * a.Cout + co);
    if (a.split > 0) {
      if (co < a.split) plane = (size_t)b * a.split + co;
      else outp = a.out1, plane = (size_t)b * (a.Cout - a.split) + (co - a.split);
      if (outp == nullptr) continue;
    }
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    const bool pair = (x + 1 < a.Wo) && ((a.Wo & 1) == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (y + i >= a.Ho) continue;
      const size_t o = (plane * a.Ho + y + i) * a.Wo + x;
      float v0 = apply_act(sr[i][0] + sr[i][1] + sr[i][2] + bv, a.act);
      float v1 = apply_act(sr[i][1] - sr[i][2] - sr[i][3] + bv, a.act);
      if (pair) {
        if (a.residual) {
          const float2 rv = __ldg(reinterpret_cast<const float2*>(a.residual + o));
          v0 += rv.x, v1 += rv.y;
        }
        *reinterpret_cast<float2*>(outp + o) = make_float2(v0, v1);
      } else {
        if (a.residual) v0 += __ldg(a.residual + o);
        outp[o] = v0;
        if (x + 1 < a.Wo) {
          if (a.residual) v1 += __ldg(a.residual + o + 1);
          outp[o + 1] = v1;
        }
      }
    }
  }
}

// U = G g G^T for every (co, ci): wt[16][Cin'][out_pad]; `transpose` builds the data-gradient operator
// (flipped taps, swapped channel roles) before transforming.
__global__ void conv_prep_wino_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int out_pad,
                                              int transpose) {
  const int n_in = transpose ? Cout : Cin;
  const int n_out = transpose ? Cin : Cout;
  const size_t total = (size_t)n_in * out_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % out_pad), rowi = (int)(i / out_pad);
    float g[3][3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v = 0.f;
        if (col < n_out) {
          if (!transpose) v = __ldg(w + ((size_t)col * Cin + rowi) * 9 + ky * 3 + kx);
          else v = __ldg(w + ((size_t)rowi * Cin + col) * 9 + (2 - ky) * 3 + (2 - kx));
        }
        g[ky][kx] = v;
      }
    float tmp[4][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      tmp[0][j] = g[0][j];
      tmp[1][j] = 0.5f * (g[0][j] + g[1][j] + g[2][j]);
      tmp[2][j] = 0.5f * (g[0][j] - g[1][j] + g[2][j]);
      tmp[3][j] = g[2][j];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float u0 = tmp[r][0], u1 = 0.5f * (tmp[r][0] + tmp[r][1] + tmp[r][2]), u2 = 0.5f * (tmp[r][0] - tmp[r][1] + tmp[r][2]),
                  u3 = tmp[r][2];
      wt[((size_t)(r * 4 + 0) * n_in + rowi) * out_pad + col] = u0;
      wt[((size_t)(r * 4 + 1) * n_in + rowi) * out_pad + col] = u1;
      wt[((size_t)(r * 4 + 2) * n_in + rowi) * out_pad + col] = u2;
      wt[((size_t)(r * 4 + 3) * n_in + rowi) * out_pad + col] = u3;
    }
  }
}

// ---- weight gradient -------------------------------------------------------------------------
constexpr int WG_GPITCH = CT_H * CT_W + 4;   // 132 (scalar variant: one channel per plane)
constexpr int WG_PPITCH = 2 * CT_H * CT_W + 4;   // 260 (packed variant: one channel PAIR per plane, pixel-major [pos][2])

struct WgradArgs {
  VirtIn vin;
  const float* g;    // (B, Cout, H, W) gradient w.r.t. the convolution output (activation already folded in)
  int B, H, W, Cin, Cout;
  int tiles_x, tiles_y;
  int items_per_split;   // (image, tile) work items per z-slice
  float* gw;         // (Cout, Cin, k, k), zero-initialised, accumulated with atomics
};

// Thread = (WCO output channels, one input channel): WCO*k*k accumulators, reduced over the (image, tile) items of
// this z-slice; CTA = 8 channel groups x 16 input channels (128 threads).  With WCO = 8 every shared-memory word feeds
// ~5.8 FMAs (g: 4 pixels x 8 channels, v: 3 x 6 taps per 288 FMAs), which keeps the kernel off the 32 words/clk
// shared-memory limit that a 4-channel thread tile (4.2 FMAs per word) sits on.
constexpr int WG_CQ = 8, WG_CI = 16;

template <int KS, int WCO>
__global__ void __launch_bounds__(CONV_THREADS) conv_wgrad_kernel(const __grid_constant__ WgradArgs a) {
  constexpr int KK = KS * KS;
  constexpr int HALO = KS / 2;
  constexpr int ROWS = CT_H + 2 * HALO, COLS = CT_W + 2 * HALO;
  constexpr int CO_T = WG_CQ * WCO;
  constexpr int NPOS = ROWS * COLS;
  constexpr int NSLOT = (NPOS + CONV_THREADS - 1) / CONV_THREADS;   // input-tile positions per thread (2)
  constexpr bool PACKED = WCO >= 2;   // channel pairs interleaved per pixel so (g[c], g[c+1]) is one 64-bit FFMA2 operand
  __shared__ __align__(16) float g_s[PACKED ? (CO_T / 2) * WG_PPITCH : CO_T * WG_GPITCH];
  __shared__ __align__(16) float v_s[WG_CI * IN_PLANE];

  const int tid = threadIdx.x;
  // warp = 8 channel groups x 4 input channels: gradient-tile loads of the 8 lanes of a quarter-warp are distinct
  // 16-byte chunks (plane stride 260 / 132 floats), input-tile loads are broadcasts.  Thread owns the output-channel
  // pairs co0 + 2*(cq + 8*c2) + {0,1} (c2 < WCO/2) [packed] or channel co0 + cq [scalar] and input channel ci0 + ci.
  const int cq = tid & (WG_CQ - 1), ci = tid >> 3;
  const int co0 = blockIdx.x * CO_T, ci0 = blockIdx.y * WG_CI;
  const int n_tiles = a.tiles_x * a.tiles_y;
  const int n_items = a.B * n_tiles;
  const int it0 = blockIdx.z * a.items_per_split;
  const int it1 = min(n_items, it0 + a.items_per_split);
  const bool vec_ok = (a.W & 3) == 0;
  const int C0 = a.vin.C0;
  const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;

  // accumulators: output-channel pairs packed for FFMA2 (WCO >= 2), scalar otherwise
  constexpr int WP2 = WCO >= 2 ? WCO / 2 : 1;
  f32x2 acc2[WP2][KK];
  float acc1[KK];
#pragma unroll
  for (int t = 0; t < KK; ++t) {
    acc1[t] = 0.f;
#pragma unroll
    for (int c = 0; c < WP2; ++c) acc2[c][t] = 0ull;
  }

  for (int it = it0; it < it1; ++it) {
    const int b = it / n_tiles, tile = it - b * n_tiles;
    const int ty0 = (tile / a.tiles_x) * CT_H, tx0 = (tile % a.tiles_x) * CT_W;
    // gradient tile: CO_T channels x 8 rows x 16 pixels, one float4 (4 pixels) per channel, thread and step
    auto load_g4 = [&](int c, int y, int x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co0 + c < a.Cout && y < a.H) {
        const float* src = a.g + (((size_t)b * a.Cout + co0 + c) * a.H + y) * a.W + x;
        if (vec_ok && x + 3 < a.W) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (x < a.W) v.x = __ldg(src);
          if (x + 1 < a.W) v.y = __ldg(src + 1);
          if (x + 2 < a.W) v.z = __ldg(src + 2);
          if (x + 3 < a.W) v.w = __ldg(src + 3);
        }
      }
      return v;
    };
    if (PACKED) {   // two channels per thread, stored pixel-interleaved: (g[c][p], g[c+1][p]) adjacent
#pragma unroll 2
      for (int i = tid; i < (CO_T / 2) * CT_H * (CT_W / 4); i += CONV_THREADS) {
        const int cp = i >> 5, rem = i & 31;
        const int r = rem >> 2, q = rem & 3;
        const float4 va = load_g4(2 * cp, ty0 + r, tx0 + 4 * q), vb = load_g4(2 * cp + 1, ty0 + r, tx0 + 4 * q);
        float4* dst = reinterpret_cast<float4*>(g_s + cp * WG_PPITCH + (r * CT_W + 4 * q) * 2);
        dst[0] = make_float4(va.x, vb.x, va.y, vb.y);
        dst[1] = make_float4(va.z, vb.z, va.w, vb.w);
      }
    } else {
      for (int i = tid; i < CO_T * CT_H * (CT_W / 4); i += CONV_THREADS) {
        const int c = i >> 5, rem = i & 31;
        const int r = rem >> 2, q = rem & 3;
        *reinterpret_cast<float4*>(g_s + c * WG_GPITCH + r * CT_W + 4 * q) = load_g4(c, ty0 + r, tx0 + 4 * q);
      }
    }
    // input tile: thread = fixed tile position(s), all 16 channels; the position's source (padding, reflection,
    // up-sampling tap) is channel independent and derived once per item
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
      const int pos = tid + sl * CONV_THREADS;
      if (pos < NPOS) {
        const int r = pos / COLS, cc = pos - r * COLS;
        TapEntry te;
        int o1;
        build_tile_map(a.vin, ty0 + r - HALO, tx0 + cc - HALO, te, o1);
        float* dst = v_s + r * IN_PITCH + cc;
        const float* x0b = a.vin.x0 + ((size_t)b * C0 + ci0) * plane0;
        const float* x1b = a.vin.x1 + ((size_t)b * a.vin.C1 + (ci0 - C0)) * plane1;
#pragma unroll
        for (int ch = 0; ch < WG_CI; ch += 8) {
          float pre[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = ci0 + ch + k;   // CTA-uniform
            float v = 0.f;
            if (c < C0) {
              if (te.o00 >= 0) {
                const float* pl = x0b + (size_t)(ch + k) * plane0;
                if (a.vin.up0 != DD_UP_BILINEAR2) {
                  v = __ldg(pl + te.o00);
                } else {
                  const float v00 = __ldg(pl + te.o00), v01 = __ldg(pl + te.o01), v10 = __ldg(pl + te.o10), v11 = __ldg(pl + te.o11);
                  v = (1.f - te.ly) * ((1.f - te.lx) * v00 + te.lx * v01) + te.ly * ((1.f - te.lx) * v10 + te.lx * v11);
                }
              }
            } else if (c < a.Cin && o1 >= 0) {
              v = __ldg(x1b + (size_t)(ch + k) * plane1 + o1);
            }
            pre[k] = v;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[(ch + k) * IN_PLANE] = pre[k];
        }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < CT_H; ++r) {
#pragma unroll
      for (int xq = 0; xq < CT_W; xq += 4) {
        float gv[4];
        f32x2 gp[WP2][4];   // packed: (g[c][p], g[c+1][p])
        if (PACKED) {
#pragma unroll
          for (int c = 0; c < WP2; ++c) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(g_s + (cq + WG_CQ * c) * WG_PPITCH + (r * CT_W + xq) * 2);
            const ulonglong2 a01 = src[0], a23 = src[1];
            gp[c][0] = a01.x, gp[c][1] = a01.y, gp[c][2] = a23.x, gp[c][3] = a23.y;
          }
        } else {
          const float4 g4 = *reinterpret_cast<const float4*>(g_s + cq * WG_GPITCH + r * CT_W + xq);
          gv[0] = g4.x, gv[1] = g4.y, gv[2] = g4.z, gv[3] = g4.w;
        }
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          const float* vp = v_s + ci * IN_PLANE + (r + dy) * IN_PITCH + xq;
          float vr[4 + KS - 1];
          const float4 v4 = *reinterpret_cast<const float4*>(vp);
          vr[0] = v4.x, vr[1] = v4.y, vr[2] = v4.z, vr[3] = v4.w;
          if (KS == 3) {
            const float2 v2 = *reinterpret_cast<const float2*>(vp + 4);
            vr[4] = v2.x, vr[5] = v2.y;
          }
          if (PACKED) {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx)
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                const f32x2 vv = pack2(vr[p + dx], vr[p + dx]);
#pragma unroll
                for (int c = 0; c < WP2; ++c) acc2[c][dy * KS + dx] = fma2(gp[c][p], vv, acc2[c][dy * KS + dx]);
              }
          } else {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx)
#pragma unroll
              for (int p = 0; p < 4; ++p) acc1[dy * KS + dx] = fmaf(gv[p], vr[p + dx], acc1[dy * KS + dx]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (ci0 + ci < a.Cin) {
#pragma unroll
    for (int c = 0; c < WCO; ++c) {
      const int co = PACKED ? co0 + 2 * (cq + WG_CQ * (c >> 1)) + (c & 1) : co0 + cq;
      if (co >= a.Cout) continue;
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        float v;
        if (PACKED) {
          float lo, hi;
          unpack2(acc2[c >> 1][t], lo, hi);
          v = (c & 1) ? hi : lo;
        } else {
          v = acc1[t];
        }
        atomicAdd(a.gw + ((size_t)co * a.Cin + ci0 + ci) * KK + t, v);
      }
    }
  }
}

// ---- weight gradient in the Winograd domain (3x3 layers with > 16 output and >= 8 input channels) ----------------
// dW = G^T [ sum over tiles (A dY A^T) (.) (B^T d B) ] G: the same F(2x2,3x3) transforms as the forward kernel, 2.25x
// fewer multiplies than the direct weight gradient.  CTA = 256 threads owns 32 output x 32 input channels and walks a
// contiguous range of K-steps; one K-step = 8 Winograd tiles (a 4x8 block of output pixels of one image).  Per step and
// frequency the product [32 co x 8 tiles] x [8 tiles x 32 ci] accumulates into thread = (frequency, 8 co, 8 ci) with
// FFMA2 on input-channel pairs — the GEMM of conv_wino_kernel with the transformed output gradient in the role of the
// weights.  Same one-barrier software pipeline: step k+2 is gathered into registers, step k+1 transformed, step k
// multiplied; half of the warps run transform and GEMM in the opposite order.
constexpr int WW_CH = 32;                      // channels per CTA on either side
constexpr int WW_KT = 8;                       // tiles per K-step: 2 tile rows x 4 tile columns
constexpr int WW_IN_R = 6, WW_IN_C = 10;       // input patch of a step (4x8 outputs + halo)
constexpr int WW_IN_PITCH = 12;
constexpr int WW_IN_PLANE = WW_IN_R * WW_IN_PITCH + 8;   // 80: two channels of a half-warp land in disjoint 64-byte quarters (conflict-free 64-bit transform loads)
constexpr int WW_G_PLANE = 40;                 // 4x8 gradient block (+8, same reason)
constexpr int WW_TP = 36;                      // pitch over channels of the transformed operands
constexpr int WW_SMEM_LOOP = 2 * WW_CH * WW_IN_PLANE + 2 * WW_CH * WW_G_PLANE + 2 * 2 * 16 * WW_KT * WW_TP;
constexpr int WW_SMEM_EPI = 16 * WW_CH * 32;
constexpr int WW_SMEM_FLOATS = WW_SMEM_LOOP > WW_SMEM_EPI ? WW_SMEM_LOOP : WW_SMEM_EPI;

struct WinoWgradArgs {
  VirtIn vin;
  const float* g;
  int B, H, W, Cin, Cout;
  int nbr, nbc;          // 4x8 pixel blocks per image (rows, columns)
  int steps_per_split;
  float* gw;
  float* gb;             // optional (Cout,) bias gradient, zero-initialised: summed by the first input-channel block
};

__global__ void __launch_bounds__(WN_THREADS, 2) conv_wgrad_wino_kernel(const __grid_constant__ WinoWgradArgs a) {
  extern __shared__ __align__(16) float wsm[];
  float* in_s = wsm;                                   // [2][32 ci][WW_IN_PLANE]
  float* g_s = in_s + 2 * WW_CH * WW_IN_PLANE;         // [2][32 co][WW_G_PLANE]
  float* V_s = g_s + 2 * WW_CH * WW_G_PLANE;           // [2][16][8 tiles][WW_TP]  transformed input (ci inner)
  float* D_s = V_s + 2 * 16 * WW_KT * WW_TP;           // [2][16][8 tiles][WW_TP]  transformed output gradient (co inner)
  float* M_s = wsm;                                    // epilogue: [16][32 co][32 ci], columns XOR-swizzled by the co group

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int co0 = blockIdx.x * WW_CH, ci0 = blockIdx.y * WW_CH;
  const int per_img = a.nbr * a.nbc;
  const int n_total = a.B * per_img;
  const int s0 = blockIdx.z * a.steps_per_split;
  const int nk = min(a.steps_per_split, n_total - s0);
  if (nk <= 0) return;
  // GEMM role: frequency p, 8 output channels (cg), 8 input channels (tg)
  const int p = 2 * warp + (lane >> 4), cg = (lane >> 2) & 3, tg = lane & 3;
  // transform role: channel tc (input channel and output channel alike), tile tt of the step
  const int tc = tid >> 3, tt = tid & 7, ttr = tt >> 2, ttc = tt & 3;
  // loader roles: input patch position ipos (60 valid) x 8 channels (quarter iq); gradient pixel gpx x 4 channels (gq)
  const int ipos = tid & 63, iq = tid >> 6;
  const int ir = ipos / WW_IN_C, ic = ipos - ir * WW_IN_C;
  const bool ipos_ok = ipos < WW_IN_R * WW_IN_C;
  const int gpx = tid & 31, gq = tid >> 5;
  const int gr = gpx >> 3, gc = gpx & 7;
  const int C0 = a.vin.C0;
  const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win, gplane = (size_t)a.H * a.W;

  f32x2 acc2[8][4];   // [output channel][input-channel pair]
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc2[c][t] = 0ull;
  const bool want_bias = a.gb != nullptr && blockIdx.y == 0;   // bias gradient rides on the transform of channel tc
  float bias_acc = 0.f;

  // K-steps are visited in order: (image, block row, block column) of the next step to gather advance incrementally
  int nb = s0 / per_img, nbr_i = (s0 - nb * per_img) / a.nbc, nbc_i = (s0 - nb * per_img) - nbr_i * a.nbc;
  auto gather = [&](float (&pin)[8], float (&pg)[4]) {
    const int b = nb, br = nbr_i, bc = nbc_i;
    if (++nbc_i == a.nbc) {
      nbc_i = 0;
      if (++nbr_i == a.nbr) nbr_i = 0, ++nb;
    }
    TapEntry te;
    int o1;
    build_tile_map(a.vin, 4 * br + ir - 1, 8 * bc + ic - 1, te, o1);
    const bool ok0 = ipos_ok && te.o00 >= 0, ok1 = ipos_ok && o1 >= 0 && a.vin.x1 != nullptr;
    const int cb = ci0 + iq * 8;
    const float* q0 = a.vin.x0 + ((size_t)b * C0 + cb) * plane0 + (ok0 ? te.o00 : 0);
    const float* q1 = ok1 ? a.vin.x1 + ((ptrdiff_t)b * a.vin.C1 + (cb - C0)) * (ptrdiff_t)plane1 + o1 : a.vin.x0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb + j;
      const bool in0 = c < C0;
      const bool ok = in0 ? ok0 : (ok1 && c < a.Cin);
      float v = 0.f;
      if (ok) v = __ldg(in0 ? q0 : q1);
      pin[j] = v;
      q0 += plane0, q1 += plane1;
    }
    const int y = 4 * br + gr, x = 8 * bc + gc;
    const bool gok = y < a.H && x < a.W;
    const float* gp = a.g + ((size_t)b * a.Cout + co0 + gq * 4) * gplane + (gok ? (size_t)y * a.W + x : 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pg[j] = (gok && co0 + gq * 4 + j < a.Cout) ? __ldg(gp) : 0.f;
      gp += gplane;
    }
  };
  auto scatter = [&](int buf, const float (&pin)[8], const float (&pg)[4]) {
    if (ipos_ok) {
      float* dst = in_s + buf * WW_CH * WW_IN_PLANE + (iq * 8) * WW_IN_PLANE + ir * WW_IN_PITCH + ic;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j * WW_IN_PLANE] = pin[j];
    }
    float* gd = g_s + buf * WW_CH * WW_G_PLANE + (gq * 4) * WW_G_PLANE + gr * 8 + gc;
#pragma unroll
    for (int j = 0; j < 4; ++j) gd[j * WW_G_PLANE] = pg[j];
  };
  // V = B^T d B of (input channel tc, tile tt) and dM = A dY A^T of (output channel tc, tile tt): buffers sb -> tb
  auto transform = [&](int sb, int tb) {
    {
      const float* dp = in_s + sb * WW_CH * WW_IN_PLANE + tc * WW_IN_PLANE + (2 * ttr) * WW_IN_PITCH + 2 * ttc;
      float d[4][4], t[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 lo = *reinterpret_cast<const float2*>(dp + i * WW_IN_PITCH);
        const float2 hi = *reinterpret_cast<const float2*>(dp + i * WW_IN_PITCH + 2);
        d[i][0] = lo.x, d[i][1] = lo.y, d[i][2] = hi.x, d[i][3] = hi.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[0][j] = d[0][j] - d[2][j];
        t[1][j] = d[1][j] + d[2][j];
        t[2][j] = d[2][j] - d[1][j];
        t[3][j] = d[1][j] - d[3][j];
      }
      float* vp = V_s + tb * 16 * WW_KT * WW_TP + tt * WW_TP + tc;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vp[(i * 4 + 0) * WW_KT * WW_TP] = t[i][0] - t[i][2];
        vp[(i * 4 + 1) * WW_KT * WW_TP] = t[i][1] + t[i][2];
        vp[(i * 4 + 2) * WW_KT * WW_TP] = t[i][2] - t[i][1];
        vp[(i * 4 + 3) * WW_KT * WW_TP] = t[i][1] - t[i][3];
      }
    }
    {
      const float* gp = g_s + sb * WW_CH * WW_G_PLANE + tc * WW_G_PLANE + (2 * ttr) * 8 + 2 * ttc;
      const float2 y0 = *reinterpret_cast<const float2*>(gp), y1 = *reinterpret_cast<const float2*>(gp + 8);
      if (want_bias) bias_acc += (y0.x + y0.y) + (y1.x + y1.y);   // every output pixel belongs to exactly one tile
      const float r[4][2] = {{y0.x, y0.y}, {y0.x + y1.x, y0.y + y1.y}, {y0.x - y1.x, y0.y - y1.y}, {-y1.x, -y1.y}};
      float* dp = D_s + tb * 16 * WW_KT * WW_TP + tt * WW_TP + tc;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dp[(i * 4 + 0) * WW_KT * WW_TP] = r[i][0];
        dp[(i * 4 + 1) * WW_KT * WW_TP] = r[i][0] + r[i][1];
        dp[(i * 4 + 2) * WW_KT * WW_TP] = r[i][0] - r[i][1];
        dp[(i * 4 + 3) * WW_KT * WW_TP] = -r[i][1];
      }
    }
  };
  auto gemm = [&](int buf) {
    const float* up = D_s + buf * 16 * WW_KT * WW_TP + p * WW_KT * WW_TP + cg * 8;
    const float* vp = V_s + buf * 16 * WW_KT * WW_TP + p * WW_KT * WW_TP + tg * 8;
#pragma unroll WINO_UNROLL
    for (int k = 0; k < WW_KT; ++k) {
      const float4 u0 = *reinterpret_cast<const float4*>(up + k * WW_TP), u1 = *reinterpret_cast<const float4*>(up + k * WW_TP + 4);
      const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(vp + k * WW_TP), v1 = *reinterpret_cast<const ulonglong2*>(vp + k * WW_TP + 4);
      const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      const f32x2 v[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const f32x2 uu = pack2(u[c], u[c]);
#pragma unroll
        for (int tq = 0; tq < 4; ++tq) acc2[c][tq] = fma2(uu, v[tq], acc2[c][tq]);
      }
    }
  };

  const bool gemm_first = (warp & 4) != 0;
  float pin[8], pg[4];
  gather(pin, pg);
  scatter(0, pin, pg);
  __syncthreads();   // staging buffer 0 visible
  if (nk > 1) gather(pin, pg);
  transform(0, 0);
  if (nk > 1) scatter(1, pin, pg);
  __syncthreads();   // transformed buffer 0, staging buffer 1 visible

  for (int k = 0; k < nk; ++k) {
    const int buf = k & 1;
    const bool next = k + 1 < nk, next2 = k + 2 < nk;
    if (next2) gather(pin, pg);   // loads in flight during the arithmetic below
    if (gemm_first) {
      gemm(buf);
      if (next) transform(buf ^ 1, buf ^ 1);
    } else {
      if (next) transform(buf ^ 1, buf ^ 1);
      gemm(buf);
    }
    if (next2) scatter(buf, pin, pg);   // staging buffer `buf` was last read by transform(k) during step k-1
    __syncthreads();
  }

  if (want_bias) {   // threads tid = 8*tc + tt: sum the 8 tile lanes of each output channel
    bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 1);
    bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 2);
    bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 4);
    if (tt == 0 && co0 + tc < a.Cout) atomicAdd(a.gb + co0 + tc, bias_acc);
  }
  // epilogue: gather the 16 frequencies of every (co, ci) through shared memory, dW = G^T dU G, accumulate atomically
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int tq = 0; tq < 4; ++tq) {
      float lo, hi;
      unpack2(acc2[c][tq], lo, hi);
      *reinterpret_cast<float2*>(M_s + (p * WW_CH + cg * 8 + c) * 32 + ((tg * 8 + 2 * tq) ^ (2 * cg))) = make_float2(lo, hi);
    }
  __syncthreads();
  for (int q = tid; q < WW_CH * WW_CH; q += WN_THREADS) {
    const int col = q >> 5, cil = q & 31;
    const int co = co0 + col, ci = ci0 + cil;
    if (co >= a.Cout || ci >= a.Cin) continue;
    float m[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) m[i][j] = M_s[((i * 4 + j) * WW_CH + col) * 32 + (cil ^ (2 * ((col >> 3) & 3)))];
    float pr[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pr[0][j] = m[0][j] + 0.5f * (m[1][j] + m[2][j]);
      pr[1][j] = 0.5f * (m[1][j] - m[2][j]);
      pr[2][j] = 0.5f * (m[1][j] + m[2][j]) + m[3][j];
    }
    float* dst = a.gw + ((size_t)co * a.Cin + ci) * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      atomicAdd(dst + i * 3 + 0, pr[i][0] + 0.5f * (pr[i][1] + pr[i][2]));
      atomicAdd(dst + i * 3 + 1, 0.5f * (pr[i][1] - pr[i][2]));
      atomicAdd(dst + i * 3 + 2, 0.5f * (pr[i][1] + pr[i][2]) + pr[i][3]);
    }
  }
}

// grad_bias[co] = sum over b, y, x of g; grid (Cout, chunks), one atomicAdd per CTA into the zeroed output
__global__ void __launch_bounds__(256) conv_bias_grad_kernel(const float* __restrict__ g, float* __restrict__ gb, int B, int Cout,
                                                             int HW) {
  __shared__ float sh[8];
  const int co = blockIdx.x;
  const size_t per_img = (size_t)HW;
  const size_t total = (size_t)B * per_img;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.y * blockDim.x) {
    const size_t b = i / per_img, p = i - b * per_img;
    acc += __ldg(g + (b * Cout + co) * per_img + p);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(gb + co, t);
  }
}

}  // namespace dd

#include "conv_wgrad_tc.cuh"  // tcgen05 weight gradient of the 3x3 layers (uses VirtIn, conv_tc4.cuh helpers)
#include "conv_pw_small.cuh"  // 1x1 layers with <= 4 output channels as 16-byte streams (uses ConvArgs, VirtIn, apply_act)

namespace dd {

// ---- host side ---------------------------------------------------------------------------------

static int device_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
      sms = 148;
  }
  return sms;
}

template <int KS>
static int wgrad_occupancy(int wco) {
  static int occ[4] = {0, 0, 0, 0};
  const int slot = wco == 8 ? 3 : (wco == 4 ? 2 : (wco == 2 ? 1 : 0));
  if (occ[slot] == 0) {
    int n = 0;
    cudaError_t e;
    if (wco == 8) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_wgrad_kernel<KS, 8>, CONV_THREADS, 0);
    else if (wco == 4) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_wgrad_kernel<KS, 4>, CONV_THREADS, 0);
    else if (wco == 2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_wgrad_kernel<KS, 2>, CONV_THREADS, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_wgrad_kernel<KS, 1>, CONV_THREADS, 0);
    occ[slot] = (e == cudaSuccess && n > 0) ? n : 4;
  }
  return occ[slot];
}

static inline int cpt_for(int cout) { return cout > 32 ? 8 : (cout > 16 ? 4 : (cout > 8 ? 2 : 1)); }
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static int validate_conv(const dd_conv_desc* d) {
  DD_REQUIRE(d != nullptr, "dd_conv_desc is NULL");
  DD_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "dd_conv: bad shape");
  DD_REQUIRE(d->ksize == 1 || d->ksize == 3, "dd_conv: ksize must be 1 or 3");
  DD_REQUIRE(d->C0 > 0 && d->x0 && d->weight, "dd_conv: x0 / weight missing");
  DD_REQUIRE(d->C1 >= 0 && (d->C1 == 0 || d->x1), "dd_conv: x1 missing");
  DD_REQUIRE(d->up0 >= 0 && d->up0 <= 2 && d->act >= 0 && d->act <= 3, "dd_conv: bad up0/act");
  DD_REQUIRE(d->up0 == DD_UP_NONE || (d->H % 2 == 0 && d->W % 2 == 0), "dd_conv: x2 up-sampling needs even H, W");
  DD_REQUIRE(!(d->residual && d->act != DD_ACT_NONE), "dd_conv: residual is only supported with DD_ACT_NONE");
  DD_REQUIRE(!(d->ksize == 3 && d->pad_mode == DD_PAD_REFLECT && (d->H < 2 || d->W < 2)), "dd_conv: reflect pad needs H, W >= 2");
  return DD_OK;
}

static VirtIn make_vin(const dd_conv_desc* d) {
  VirtIn v;
  v.x0 = d->x0, v.x1 = d->x1, v.C0 = d->C0, v.C1 = d->C1;
  v.up0 = d->up0;
  v.H0 = d->up0 == DD_UP_NONE ? d->H : d->H / 2;
  v.W0 = d->up0 == DD_UP_NONE ? d->W : d->W / 2;
  v.Hin = d->H, v.Win = d->W;
  v.pad_mode = d->ksize == 3 ? d->pad_mode : DD_PAD_ZERO;
  return v;
}

template <int KS>
static void launch_core(const ConvArgs& args, int cpt, dim3 grid_base, cudaStream_t st) {
  dim3 grid = grid_base;
  grid.y = (args.Cout + 8 * cpt - 1) / (8 * cpt);
  if (cpt == 8) { conv_core_kernel<KS, 8><<<grid, CONV_THREADS, 0, st>>>(args); dd::count_launches(1); }
  else if (cpt == 4) { conv_core_kernel<KS, 4><<<grid, CONV_THREADS, 0, st>>>(args); dd::count_launches(1); }
  else if (cpt == 2) { conv_core_kernel<KS, 2><<<grid, CONV_THREADS, 0, st>>>(args); dd::count_launches(1); }
  else { conv_core_kernel<KS, 1><<<grid, CONV_THREADS, 0, st>>>(args); dd::count_launches(1); }
}

static bool use_winograd(int ks, int cin, int cout) {
  static const bool disabled = getenv("DD_NO_WINOGRAD") != nullptr;
  return !disabled && ks == 3 && cout > 16 && cin >= 8;
}

static int run_core(ConvArgs& args, int ks, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose,
                    cudaStream_t st) {
  if (use_pw_small(ks, transpose ? args.Cin : args.Cout, args.Ho, args.Wo, args.vin.up0))   // motion-decoder 1x1 reductions
    return transpose ? run_pw_small_dgrad(args, wt_buf, w_oihw, st) : run_pw_small_fwd(args, wt_buf, w_oihw, st);
  if (use_tc4_conv(ks, args.Cin, args.Cout))  // tensor cores (3xTF32, fp32 accuracy): every 3x3 layer with more than 16 output channels
    return run_conv_tc4(args, wt_buf, w_oihw, Cout_f, Cin_f, transpose, device_sms(), st);
  if (use_winograd(ks, args.Cin, args.Cout)) {
    args.cout_pad = round_up(args.Cout, WN_CO);
    const size_t wn = (size_t)args.Cin * args.cout_pad;
    conv_prep_wino_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(
        w_oihw, wt_buf, Cout_f, Cin_f, args.cout_pad, transpose ? 1 : 0);
    dd::count_launches(1);
    args.wt = wt_buf;
    args.tiles_x = (args.Wo + CT_W - 1) / CT_W;
    const int tiles_y = (args.Ho + CT_H - 1) / CT_H;
    static bool configured = false;
    const size_t smem = WN_SMEM_FLOATS * sizeof(float);
    if (!configured) {
      DD_CHECK_CUDA(cudaFuncSetAttribute(conv_wino_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_wino_kernel: bilinear up-sampling must be materialised first");
    dim3 grid(args.tiles_x * tiles_y, args.cout_pad / WN_CO, args.B);
    conv_wino_kernel<<<grid, WN_THREADS, smem, st>>>(args);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  const int KK = ks * ks;
  static const bool no_small = getenv("DD_NO_SMALL_CONV") != nullptr;
  if (!no_small && ks == 3 && args.Cout <= 16 && args.vin.up0 == DD_UP_NONE) {
    const int co_t = args.Cout <= 2 ? 2 : (args.Cout <= 4 ? 4 : (args.Cout <= 12 ? 12 : 16));
    args.cout_pad = co_t;
    const size_t wn = (size_t)args.Cin * KK * co_t;
    conv_prep_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(
        w_oihw, wt_buf, Cout_f, Cin_f, KK, co_t, transpose ? 1 : 0); dd::count_launches(1);
    args.wt = wt_buf;
    args.tiles_x = (args.Wo + SM_TW - 1) / SM_TW;
    const int tiles_y = (args.Ho + SM_TH - 1) / SM_TH;
    dim3 grid(args.tiles_x * tiles_y, 1, args.B);
    if (co_t == 2) conv_small_kernel<2><<<grid, SM_THREADS, 0, st>>>(args);
    else if (co_t == 4) conv_small_kernel<4><<<grid, SM_THREADS, 0, st>>>(args);
    else if (co_t == 12) conv_small_kernel<12><<<grid, SM_THREADS, 0, st>>>(args);
    else conv_small_kernel<16><<<grid, SM_THREADS, 0, st>>>(args);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  const int cpt = cpt_for(args.Cout);
  args.cout_pad = round_up(args.Cout, 8 * cpt);
  const size_t wn = (size_t)args.Cin * KK * args.cout_pad;
  conv_prep_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(
      w_oihw, wt_buf, Cout_f, Cin_f, KK, args.cout_pad, transpose ? 1 : 0); dd::count_launches(1);
  args.wt = wt_buf;
  args.tiles_x = (args.Wo + CT_W - 1) / CT_W;
  const int tiles_y = (args.Ho + CT_H - 1) / CT_H;
  dim3 grid(args.tiles_x * tiles_y, 1, args.B);
  if (ks == 3) launch_core<3>(args, cpt, grid, st);
  else launch_core<1>(args, cpt, grid, st);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

struct ConvWs {
  size_t wt, up, gconv, wtd, gpad, slabs, total;
  size_t fwd_bytes;   // what the forward pass needs (prepared weights + materialised up-sampling)
};

static void launch_resize_fwd(const float* x, float* out, int BC, int hi, int wi, int ho, int wo, int sigmoid, cudaStream_t st);

// Bilinear x2 up-sampling of x0 (depth_decoder.py:104) is materialised once per call into the workspace (a few tens of
// MB, written and read once at HBM speed) so that the convolution loaders stay one tap per element; returns the
// virtual-input description the kernels should use.
static VirtIn materialise_up(const dd_conv_desc* d, void* workspace, const ConvWs& ws, cudaStream_t st) {
  VirtIn v = make_vin(d);
  if (d->up0 != DD_UP_BILINEAR2) return v;
  float* up = reinterpret_cast<float*>((char*)workspace + ws.up);
  launch_resize_fwd(d->x0, up, d->B * d->C0, d->H / 2, d->W / 2, d->H, d->W, 0, st);
  v.x0 = up, v.H0 = d->H, v.W0 = d->W, v.up0 = DD_UP_NONE;
  return v;
}

static ConvWs conv_ws(const dd_conv_desc* d) {
  ConvWs w;
  const int KK = d->ksize * d->ksize;
  const int Cin = d->C0 + d->C1;
  // prepared weights: the largest of the Winograd layout [16][Cin][cout_pad] and the tensor-core layout [Cout][KK][cin_pad]
  size_t wt_f = (size_t)round_up(Cin, 32) * (KK == 9 ? 16 : KK) * round_up(d->Cout, 32) * sizeof(float);
  if (KK == 9) {   // tensor-core core: pre-split hi / lo weight blocks of the forward and of the transposed (data-gradient) problem
    const size_t f = conv_tc4_weight_bytes(Cin, d->Cout), t = conv_tc4_weight_bytes(d->Cout, Cin);
    wt_f = wt_f > f ? wt_f : f;
    wt_f = wt_f > t ? wt_f : t;
  }
  const size_t wt_d = wt_f;
  const bool reflect = d->ksize == 3 && d->pad_mode == DD_PAD_REFLECT;
  const size_t Hp = d->H + (reflect ? 2 : 0), Wp = d->W + (reflect ? 2 : 0);
  w.wt = 0;
  w.up = align256(wt_f);
  w.gconv = w.up + (d->up0 == DD_UP_BILINEAR2 ? align256((size_t)d->B * d->C0 * d->H * d->W * sizeof(float)) : 0);
  w.fwd_bytes = w.gconv;
  w.wtd = w.gconv + align256((size_t)d->B * d->Cout * d->H * d->W * sizeof(float));
  w.gpad = w.wtd + align256(wt_d);
  w.slabs = w.gpad + align256((size_t)d->B * Cin * Hp * Wp * sizeof(float));
  w.total = w.slabs + (use_tc_wgrad(d->ksize, Cin, d->Cout) ? align256(conv_wgrad_tc_slab_bytes(d->B, d->H, d->W, Cin, d->Cout, device_sms())) : 0);
  if (use_pw_small(d->ksize, d->Cout, d->H, d->W, d->up0)) w.total = w.slabs + align256(pw_small_wgrad_bytes(d->B, d->H, d->W, Cin));
  return w;
}

int conv_fwd_impl(const dd_conv_desc* d, float* out, void* workspace, size_t bytes, cudaStream_t st) {
  int rc = validate_conv(d);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(out != nullptr, "dd_conv_fwd: out is NULL");
  const ConvWs ws = conv_ws(d);
  if (!workspace || bytes < ws.fwd_bytes) {
    set_error("dd_conv_fwd: workspace too small (%zu < %zu)", bytes, ws.fwd_bytes);
    return DD_ERR_WORKSPACE;
  }
  ConvArgs args;
  memset(&args, 0, sizeof(args));
  args.vin = materialise_up(d, workspace, ws, st);
  args.B = d->B, args.Ho = d->H, args.Wo = d->W, args.oy = 0, args.ox = 0;
  args.Cin = d->C0 + d->C1, args.Cout = d->Cout;
  args.bias = d->bias, args.residual = d->residual, args.act = d->act, args.out = out;
  return run_core(args, d->ksize, reinterpret_cast<float*>((char*)workspace + ws.wt), d->weight, d->Cout, args.Cin, false, st);
}

int conv_bwd_impl(const dd_conv_desc* d, const float* out, const float* grad_out, float* grad_x0, float* grad_x1,
                  float* grad_weight, float* grad_bias, void* workspace, size_t bytes, cudaStream_t st) {
  int rc = validate_conv(d);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(grad_out != nullptr, "dd_conv_bwd: grad_out is NULL");
  DD_REQUIRE(d->act == DD_ACT_NONE || out != nullptr, "dd_conv_bwd: forward output required for the activation derivative");
  const ConvWs ws = conv_ws(d);
  if (!workspace || bytes < ws.total) {
    set_error("dd_conv_bwd: workspace too small (%zu < %zu)", bytes, ws.total);
    return DD_ERR_WORKSPACE;
  }
  const int Cin = d->C0 + d->C1, KK = d->ksize * d->ksize;
  const size_t n_out = (size_t)d->B * d->Cout * d->H * d->W;
  const float* g = grad_out;
  if (d->act != DD_ACT_NONE) {
    float* gc = reinterpret_cast<float*>((char*)workspace + ws.gconv);
    conv_act_grad_kernel<<<(int)((n_out + 255) / 256 < 2368 ? (n_out + 255) / 256 : 2368), 256, 0, st>>>(grad_out, out, gc, n_out, d->act); dd::count_launches(1);
    g = gc;
  }
  static const bool no_wino_wgrad = getenv("DD_NO_WINO_WGRAD") != nullptr;
  const bool tc_wgrad = grad_weight && use_tc_wgrad(d->ksize, Cin, d->Cout);
  const bool wino_wgrad = grad_weight && !tc_wgrad && !no_wino_wgrad && use_winograd(d->ksize, Cin, d->Cout);
  if (grad_bias) DD_CHECK_CUDA(cudaMemsetAsync(grad_bias, 0, (size_t)d->Cout * sizeof(float), st));
  if (grad_bias && !wino_wgrad && !tc_wgrad) {   // (the Winograd / tensor-core weight-gradient kernels sum the bias gradient on the way)
    const size_t total = (size_t)d->B * d->H * d->W;
    int chunks = (int)((total + 256 * 8 - 1) / (256 * 8));              // >= 8 elements per thread
    const int cap = (148 * 8 + d->Cout - 1) / d->Cout;                   // ~8 CTAs per SM over all channels
    chunks = chunks < 1 ? 1 : (chunks > cap ? cap : chunks);
    conv_bias_grad_kernel<<<dim3(d->Cout, chunks), 256, 0, st>>>(g, grad_bias, d->B, d->Cout, d->H * d->W);
    dd::count_launches(1);
  }
  if (tc_wgrad) {   // tensor cores: partial sums per work item in workspace slabs, deterministic reduction (no memset, no atomics)
    WgradTcArgs wt;
    memset(&wt, 0, sizeof(wt));
    wt.vin = materialise_up(d, workspace, ws, st);
    wt.g = g, wt.B = d->B, wt.H = d->H, wt.W = d->W, wt.Cin = Cin, wt.Cout = d->Cout;
    wt.slabs = reinterpret_cast<float*>((char*)workspace + ws.slabs);
    wt.gb = grad_bias;
    rc = run_conv_wgrad_tc(wt, grad_weight, device_sms(), st);
    if (rc != DD_OK) return rc;
  } else if (grad_weight && use_pw_small(d->ksize, d->Cout, d->H, d->W, d->up0)) {
    rc = run_pw_small_wgrad(make_vin(d), g, d->B, d->H, d->W, Cin, d->Cout, reinterpret_cast<float*>((char*)workspace + ws.slabs), grad_weight, st);
    if (rc != DD_OK) return rc;
  } else if (grad_weight) {
    DD_CHECK_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)d->Cout * Cin * KK * sizeof(float), st));
    if (wino_wgrad) {
      WinoWgradArgs ww;
      memset(&ww, 0, sizeof(ww));
      ww.vin = materialise_up(d, workspace, ws, st);
      ww.g = g, ww.B = d->B, ww.H = d->H, ww.W = d->W, ww.Cin = Cin, ww.Cout = d->Cout, ww.gw = grad_weight, ww.gb = grad_bias;
      ww.nbr = (d->H + 3) / 4, ww.nbc = (d->W + 7) / 8;
      const int n_steps = d->B * ww.nbr * ww.nbc;
      const int gx = (d->Cout + WW_CH - 1) / WW_CH, gy = (Cin + WW_CH - 1) / WW_CH;
      static bool configured = false;
      const size_t smem = WW_SMEM_FLOATS * sizeof(float);
      if (!configured) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_wino_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      int splits = (device_sms() * 2) / (gx * gy);   // one wave of the two resident CTAs per SM
      splits = splits < 1 ? 1 : (splits > n_steps ? n_steps : splits);
      ww.steps_per_split = (n_steps + splits - 1) / splits;
      splits = (n_steps + ww.steps_per_split - 1) / ww.steps_per_split;
      conv_wgrad_wino_kernel<<<dim3(gx, gy, splits), WN_THREADS, smem, st>>>(ww);
      dd::count_launches(1);
      DD_CHECK_CUDA(cudaGetLastError());
    } else {
    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.vin = materialise_up(d, workspace, ws, st);
    wa.g = g, wa.B = d->B, wa.H = d->H, wa.W = d->W, wa.Cin = Cin, wa.Cout = d->Cout, wa.gw = grad_weight;
    wa.tiles_x = (d->W + CT_W - 1) / CT_W, wa.tiles_y = (d->H + CT_H - 1) / CT_H;
    const int n_items = d->B * wa.tiles_x * wa.tiles_y;
    const int wco = d->Cout > 32 ? 8 : (d->Cout > 16 ? 4 : (d->Cout > 8 ? 2 : 1));
    const int gx = (d->Cout + WG_CQ * wco - 1) / (WG_CQ * wco), gy = (Cin + WG_CI - 1) / WG_CI;
    // exactly one wave of resident CTAs (a partial second wave would leave most SMs idle for its whole duration)
    const int resident = device_sms() * (d->ksize == 3 ? wgrad_occupancy<3>(wco) : wgrad_occupancy<1>(wco));
    int splits = resident / (gx * gy);
    splits = splits < 1 ? 1 : (splits > n_items ? n_items : splits);
    wa.items_per_split = (n_items + splits - 1) / splits;
    splits = (n_items + wa.items_per_split - 1) / wa.items_per_split;
    dim3 grid(gx, gy, splits);
#define DD_WGRAD(KSZ)                                                                    \
  do {                                                                                   \
    if (wco == 8) conv_wgrad_kernel<KSZ, 8><<<grid, CONV_THREADS, 0, st>>>(wa);          \
    else if (wco == 4) conv_wgrad_kernel<KSZ, 4><<<grid, CONV_THREADS, 0, st>>>(wa);     \
    else if (wco == 2) conv_wgrad_kernel<KSZ, 2><<<grid, CONV_THREADS, 0, st>>>(wa);     \
    else conv_wgrad_kernel<KSZ, 1><<<grid, CONV_THREADS, 0, st>>>(wa);                   \
    dd::count_launches(1);                                                               \
  } while (0)
    if (d->ksize == 3) DD_WGRAD(3); else DD_WGRAD(1);
#undef DD_WGRAD
    }
  }
  if (grad_x0 || grad_x1) {
    const bool reflect = d->ksize == 3 && d->pad_mode == DD_PAD_REFLECT;
    float* gpad = reinterpret_cast<float*>((char*)workspace + ws.gpad);
    ConvArgs args;
    memset(&args, 0, sizeof(args));
    args.vin.x0 = g, args.vin.x1 = nullptr, args.vin.C0 = d->Cout, args.vin.C1 = 0;
    args.vin.H0 = d->H, args.vin.W0 = d->W, args.vin.up0 = DD_UP_NONE, args.vin.Hin = d->H, args.vin.Win = d->W;
    args.vin.pad_mode = DD_PAD_ZERO;
    args.B = d->B;
    args.Ho = d->H + (reflect ? 2 : 0), args.Wo = d->W + (reflect ? 2 : 0);
    args.oy = reflect ? -1 : 0, args.ox = reflect ? -1 : 0;
    args.Cin = d->Cout, args.Cout = Cin;
    args.act = DD_ACT_NONE, args.out = gpad;
    // zero padding (or 1x1) without up-sampling: the padded-grid gradient IS the input gradient, so the core
    // writes grad_x0 / grad_x1 directly and the routing pass (and its round trip through HBM) disappears
    const bool direct = !reflect && d->up0 == DD_UP_NONE;
    if (direct) {
      if (d->C1 > 0) {
        args.out = grad_x0, args.out1 = grad_x1, args.split = d->C0;
        if (grad_x0 == nullptr) {   // only the skip gradient is wanted: keep the kernel's primary pointer valid
          args.out = gpad;          // scratch for the first C0 channels
        }
      } else {
        args.out = grad_x0;
      }
    }
    rc = run_core(args, d->ksize, reinterpret_cast<float*>((char*)workspace + ws.wtd), d->weight, d->Cout, Cin, true, st);
    if (rc != DD_OK) return rc;
    if (direct) {
      DD_CHECK_CUDA(cudaGetLastError());
      return DD_OK;
    }
    RouteArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.gpad = gpad, ra.B = d->B, ra.Cin = Cin, ra.H = d->H, ra.W = d->W;
    ra.Hp = args.Ho, ra.Wp = args.Wo, ra.off = reflect ? 1 : 0, ra.reflect = reflect ? 1 : 0;
    ra.C0 = d->C0, ra.C1 = d->C1, ra.up0 = d->up0;
    ra.H0 = d->up0 == DD_UP_NONE ? d->H : d->H / 2, ra.W0 = d->up0 == DD_UP_NONE ? d->W : d->W / 2;
    ra.gx0 = grad_x0, ra.gx1 = d->C1 > 0 ? grad_x1 : nullptr;
    const size_t n = (grad_x0 ? (size_t)d->B * d->C0 * ra.H0 * ra.W0 : 0) + (ra.gx1 ? (size_t)d->B * d->C1 * d->H * d->W : 0);
    if (n > 0) { conv_route_kernel<<<(int)((n + 255) / 256 < 2368 ? (n + 255) / 256 : 2368), 256, 0, st>>>(ra); dd::count_launches(1); }
  }
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

// ---- bilinear resize (align_corners=False), general sizes -----------------------------------------
__device__ __forceinline__ void resize_taps(int dst, int n_in, int n_out, int& i0, int& i1, float& l) {
  if (n_in == n_out) {
    i0 = i1 = dst, l = 0.f;
    return;
  }
  const float scale = (float)n_in / (float)n_out;
  const float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)src, n_in - 1);
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l = src - (float)i0;
}

__global__ void resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int BC, int hi, int wi, int ho, int wo,
                                  int sigmoid) {
  const size_t n = (size_t)BC * ho * wo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % wo), yo = (int)((i / wo) % ho);
    const size_t bc = i / ((size_t)wo * ho);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_taps(yo, hi, ho, y0, y1, ly);
    resize_taps(xo, wi, wo, x0, x1, lx);
    const float* p = x + bc * hi * wi;
    float v = (1.f - ly) * ((1.f - lx) * __ldg(p + y0 * wi + x0) + lx * __ldg(p + y0 * wi + x1)) +
              ly * ((1.f - lx) * __ldg(p + y1 * wi + x0) + lx * __ldg(p + y1 * wi + x1));
    if (sigmoid) v = 1.f / (1.f + expf(-v));
    out[i] = v;
  }
}

// Exact x2 up-sampling (every bilinear resize of the decoders: depth_decoder.py:104,:112-113, motion_decoder.py:38 from one level to
// the next): thread = 2 output rows x 4 output columns from a 3 x 4 input neighbourhood (12 loads for 8 outputs instead of 32, one
// 16-byte store per row).  Taps, weights and the order of operations are those of resize_taps / resize_fwd_kernel for scale 1/2
// (src = dst / 2 - 1/4 clamped at 0), so the results are bit-identical to the generic kernel.
__global__ void __launch_bounds__(256) resize2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int BC, int hi, int wi, int sigmoid) {
  const int ho = 2 * hi, wo = 2 * wi, wq = wo >> 2;   // wq output quads per row (wi even)
  const size_t n = (size_t)BC * hi * wq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % wq), k = (int)((i / wq) % hi);
    const size_t bc = i / ((size_t)wq * hi);
    const float* p = x + bc * hi * wi;
    // input rows k-1, k, k+1 and columns 2j-1 .. 2j+2 (clamped); v[r][c]
    const int rows[3] = {max(k - 1, 0), k, min(k + 1, hi - 1)};
    const int cols[4] = {max(2 * j - 1, 0), 2 * j, 2 * j + 1, min(2 * j + 2, wi - 1)};
    float v[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) v[r][c] = __ldg(p + rows[r] * wi + cols[c]);
    // per output column m of the quad: (index of tap 0, index of tap 1, weight of tap 1) into cols[]
    int ca[4] = {0, 1, 1, 2}, cb[4] = {1, 2, 2, 3};
    float lx[4] = {0.75f, 0.25f, 0.75f, 0.25f};
    if (j == 0) ca[0] = 1, cb[0] = 2, lx[0] = 0.f;   // dst = 0: src clamps to 0 -> taps (0, 1), weight 0
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // output row 2k + half: taps (k-1, k; 3/4) or (k, k+1; 1/4); dst = 0: taps (0, 1), weight 0
      int ra = half == 0 ? 0 : 1, rb = half == 0 ? 1 : 2;
      float ly = half == 0 ? 0.75f : 0.25f;
      if (half == 0 && k == 0) ra = 1, rb = 2, ly = 0.f;
      float o[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float v00 = v[ra][ca[m]], v01 = v[ra][cb[m]], v10 = v[rb][ca[m]], v11 = v[rb][cb[m]];
        o[m] = (1.f - ly) * ((1.f - lx[m]) * v00 + lx[m] * v01) + ly * ((1.f - lx[m]) * v10 + lx[m] * v11);
        if (sigmoid) o[m] = 1.f / (1.f + expf(-o[m]));
      }
      *reinterpret_cast<float4*>(out + (bc * ho + 2 * k + half) * wo + 4 * j) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

static void launch_resize_fwd(const float* x, float* out, int BC, int hi, int wi, int ho, int wo, int sigmoid, cudaStream_t st) {
  if (ho == 2 * hi && wo == 2 * wi && wi % 2 == 0 && wi >= 2 && hi >= 2 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const size_t n = (size_t)BC * hi * (wo / 4);
    resize2x_fwd_kernel<<<(int)((n + 255) / 256 < 2368 ? (n + 255) / 256 : 2368), 256, 0, st>>>(x, out, BC, hi, wi, sigmoid);
  } else {
    const size_t n = (size_t)BC * ho * wo;
    resize_fwd_kernel<<<(int)((n + 255) / 256 < 2368 ? (n + 255) / 256 : 2368), 256, 0, st>>>(x, out, BC, hi, wi, ho, wo, sigmoid);
  }
  dd::count_launches(1);
}

__global__ void resize_bwd_kernel(const float* __restrict__ go, const float* __restrict__ out, float* __restrict__ gx, int BC,
                                  int hi, int wi, int ho, int wo, int sigmoid) {
  const size_t n = (size_t)BC * ho * wo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % wo), yo = (int)((i / wo) % ho);
    const size_t bc = i / ((size_t)wo * ho);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_taps(yo, hi, ho, y0, y1, ly);
    resize_taps(xo, wi, wo, x0, x1, lx);
    float g = __ldg(go + i);
    if (sigmoid) {
      const float o = __ldg(out + i);
      g *= o * (1.f - o);
    }
    float* p = gx + bc * hi * wi;
    atomicAdd(p + y0 * wi + x0, g * (1.f - ly) * (1.f - lx));
    atomicAdd(p + y0 * wi + x1, g * (1.f - ly) * lx);
    atomicAdd(p + y1 * wi + x0, g * ly * (1.f - lx));
    atomicAdd(p + y1 * wi + x1, g * ly * lx);
  }
}

}  // namespace dd

extern "C" {

size_t dd_conv_workspace_bytes(const dd_conv_desc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0 || d->C0 <= 0 || (d->ksize != 1 && d->ksize != 3)) return 0;
  return dd::conv_ws(d).total;
}

int dd_conv_fwd(const dd_conv_desc* desc, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  return dd::conv_fwd_impl(desc, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dd_conv_bwd(const dd_conv_desc* desc, const float* out, const float* grad_out, float* grad_x0, float* grad_x1,
                float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes, void* stream) {
  return dd::conv_bwd_impl(desc, out, grad_out, grad_x0, grad_x1, grad_weight, grad_bias, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int dd_resize_bilinear_fwd(const float* x, int BC, int h_in, int w_in, int h_out, int w_out, int sigmoid, float* out,
                           void* stream) {
  using namespace dd;
  DD_REQUIRE(x && out && BC > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "dd_resize_bilinear_fwd: bad arguments");
  launch_resize_fwd(x, out, BC, h_in, w_in, h_out, w_out, sigmoid, (cudaStream_t)stream);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_resize_bilinear_bwd(const float* grad_out, const float* out, int BC, int h_in, int w_in, int h_out, int w_out,
                           int sigmoid, float* grad_x, void* stream) {
  using namespace dd;
  DD_REQUIRE(grad_out && grad_x && BC > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "dd_resize_bilinear_bwd: bad arguments");
  DD_REQUIRE(!sigmoid || out, "dd_resize_bilinear_bwd: forward output required with sigmoid");
  cudaStream_t st = (cudaStream_t)stream;
  DD_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)BC * h_in * w_in * sizeof(float), st));
  const size_t n = (size_t)BC * h_out * w_out;
  resize_bwd_kernel<<<(int)((n + 255) / 256 < 2368 ? (n + 255) / 256 : 2368), 256, 0, st>>>(grad_out, out, grad_x, BC, h_in,
                                                                                            w_in, h_out, w_out, sigmoid); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
