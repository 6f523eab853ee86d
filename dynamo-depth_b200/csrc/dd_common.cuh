// Shared device/host helpers for libdynamo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dynamo_b200.h"

namespace dd {

void set_error(const char* fmt, ...);
void count_launches(int n);   // bookkeeping for dd_launch_count() (bench.py reports it)

#define DD_CHECK_CUDA(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      dd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DD_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define DD_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      dd::set_error(__VA_ARGS__);  \
      return DD_ERR_INVALID;       \
    }                              \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Packed fp32 pairs (sm_100a FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per issue slot, each lane
// rounded exactly like the scalar instruction).  A pair lives in one 64-bit register (lo = first element).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 1/x for normal, finite x: MUFU.RCP + one Newton step (two FMAs).  The result is the correctly rounded reciprocal for
// all but a vanishing fraction of inputs (and within 1 ulp for those) without __frcp_rn's range check and slow path;
// the quantities inverted here (depth scale, projective z + 1e-7, SSIM denominators >= C1*C2) are never denormal or inf.
__device__ __forceinline__ float rcp_nr(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  const float e = fmaf(-x, r, 1.f);
  return fmaf(r, e, r);
}

// Bilinear up-sampling taps of F.interpolate(mode='bilinear', align_corners=False) from an axis of
// n_in = n_out >> shift samples (ATen UpSample.h area_pixel_compute_source_index): src =
// (dst+0.5)/2^shift - 0.5 clamped at 0, second tap min(i0+1, n_in-1).  shift==0 is a plain copy.
struct Taps {
  int i0, i1;
  float l;
};

__device__ __forceinline__ Taps up_taps(int dst, int shift, int n_in) {
  Taps t;
  if (shift == 0) {
    t.i0 = dst;
    t.i1 = dst;
    t.l = 0.f;
    return t;
  }
  const float scale = 1.f / (float)(1 << shift);
  const float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  t.i0 = (int)src;
  t.i1 = min(t.i0 + 1, n_in - 1);
  t.l = src - (float)t.i0;
  return t;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ p, int w, const Taps& ty, const Taps& tx) {
  if (ty.i0 == ty.i1 && tx.i0 == tx.i1) return __ldg(p + ty.i0 * w + tx.i0);   // same-size level: plain copy
  const float v00 = __ldg(p + ty.i0 * w + tx.i0), v01 = __ldg(p + ty.i0 * w + tx.i1);
  const float v10 = __ldg(p + ty.i1 * w + tx.i0), v11 = __ldg(p + ty.i1 * w + tx.i1);
  return (1.f - ty.l) * ((1.f - tx.l) * v00 + tx.l * v01) + ty.l * ((1.f - tx.l) * v10 + tx.l * v11);
}

// weight with which up-sampled sample `dst` reads low-res sample `i` (transpose of up_taps)
__device__ __forceinline__ float up_weight(int dst, int shift, int n_in, int i) {
  const Taps t = up_taps(dst, shift, n_in);
  float w = 0.f;
  if (t.i0 == i) w += 1.f - t.l;
  if (t.i1 == i && shift != 0) w += t.l;
  return w;
}

// ReflectionPad2d(1) index map for positions -1..n
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

}  // namespace dd
