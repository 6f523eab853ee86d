// 1x1 convolutions with at most four output channels: the `refine_motion_redu` layers of the motion decoder
// (reference networks/motion_decoder.py:30-31,:58: nn.Conv2d(2c, out_dim, 1) over cat(x1, x2) + residual; out_dim = 3 flow /
// 1 mask) -- forward, data gradient and weight gradient.  These layers are pure streams (a 128-channel input at 1/2
// resolution is 500 MB at bs32, 2 FLOP per byte); the generic direct core ran them at 1/3 .. 1/7 of the HBM rate because
// its 8x16-pixel x 8-channel mapping leaves most lanes without work when Cout <= 4.  Here:
//   forward   CTA = 8 warps = 8 channel groups over the same 128 pixels; lane = 4 consecutive pixels (16-byte loads);
//             every warp walks the channels g, g+8, ... with CO x 4 running sums; the eight partial sums meet in shared
//             memory; bias, activation and residual in the final pass
//   data grad the same mapping transposed: a warp loads the CO gradient vectors of its pixels once and writes one
//             16-byte vector per input channel (pure write stream), straight into grad_x0 / grad_x1
//   weight    warp = one input channel x a chunk of 4096 pixels of one image, CO dot products per lane, shuffle reduction,
//   grad      partials [chunk][Cin][CO] and a fixed-order second stage (deterministic; no memset, no atomics)
// Included by conv.cu (uses ConvArgs / apply_act).
#pragma once

namespace dd {

constexpr int PW_THREADS = 256;
constexpr int PW_GROUPS = PW_THREADS / 32;
constexpr int PW_CHUNK_Q = 1024;   // 16-byte pixel vectors per weight-gradient work item

__device__ __forceinline__ const float* pw_plane(const VirtIn& v, int b, int ci, size_t HW) {
  return ci < v.C0 ? v.x0 + ((size_t)b * v.C0 + ci) * HW : v.x1 + ((size_t)b * v.C1 + (ci - v.C0)) * HW;
}

// a.wt = weight table [Cin][4] (output channels padded to four, conv_prep_weights_kernel): ONE 16-byte load per input channel
// brings all output-channel weights.  SPLIT: the CTA's 8 warps share 128 pixels and split the input channels (many channels);
// otherwise every warp owns its own 128 pixels and walks all channels (few channels: no cross-warp reduction).
template <int CO, bool SPLIT>
__global__ void __launch_bounds__(PW_THREADS) pw_small_fwd_kernel(const __grid_constant__ ConvArgs a) {
  __shared__ float4 part[SPLIT ? PW_GROUPS : 1][CO][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, b = blockIdx.z;
  const size_t HW = (size_t)a.Ho * a.Wo;
  const int HW4 = (int)(HW >> 2);
  const int q = SPLIT ? blockIdx.x * 32 + lane : (blockIdx.x * PW_GROUPS + grp) * 32 + lane;
  const bool live = q < HW4;
  float4 acc[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) acc[co] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
#pragma unroll 4
    for (int ci = SPLIT ? grp : 0; ci < a.Cin; ci += SPLIT ? PW_GROUPS : 1) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pw_plane(a.vin, b, ci, HW)) + q);
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wt) + ci);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int co = 0; co < CO; ++co) {
        acc[co].x = fmaf(w[co], v.x, acc[co].x), acc[co].y = fmaf(w[co], v.y, acc[co].y);
        acc[co].z = fmaf(w[co], v.z, acc[co].z), acc[co].w = fmaf(w[co], v.w, acc[co].w);
      }
    }
  }
  auto finish = [&](int co, int qq, float4 s) {
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    s = make_float4(apply_act(s.x + bv, a.act), apply_act(s.y + bv, a.act), apply_act(s.z + bv, a.act), apply_act(s.w + bv, a.act));
    const size_t o = ((size_t)b * a.Cout + co) * HW + (size_t)qq * 4;
    if (a.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(a.residual + o));
      s.x += r.x, s.y += r.y, s.z += r.z, s.w += r.w;
    }
    *reinterpret_cast<float4*>(a.out + o) = s;
  };
  if (!SPLIT) {
    if (live) {
#pragma unroll
      for (int co = 0; co < CO; ++co)
        if (co < a.Cout) finish(co, q, acc[co]);
    }
    return;
  }
#pragma unroll
  for (int co = 0; co < CO; ++co) part[grp][co][lane] = acc[co];
  __syncthreads();
  for (int i = threadIdx.x; i < CO * 32; i += PW_THREADS) {
    const int co = i >> 5, l = i & 31, qq = blockIdx.x * 32 + l;
    if (co >= a.Cout || qq >= HW4) continue;
    float4 s = part[0][co][l];
#pragma unroll
    for (int g = 1; g < PW_GROUPS; ++g) {
      const float4 p = part[g][co][l];
      s.x += p.x, s.y += p.y, s.z += p.z, s.w += p.w;
    }
    finish(co, qq, s);
  }
}

// Data gradient: a.vin.x0 = g (B, CO_real = a.Cin, HW); a.Cout = channels of the layer's input; a.wt = the forward table
// [a.Cout][4]; channels < split go to a.out (split channels per image), the rest to a.out1; a NULL destination is skipped.
template <int CO>
__global__ void __launch_bounds__(PW_THREADS) pw_small_dgrad_kernel(const __grid_constant__ ConvArgs a) {
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, b = blockIdx.z;
  const size_t HW = (size_t)a.Ho * a.Wo;
  const int HW4 = (int)(HW >> 2);
  const int q = blockIdx.x * 32 + lane;
  if (q >= HW4) return;
  float4 g[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co)
    g[co] = co < a.Cin ? __ldg(reinterpret_cast<const float4*>(a.vin.x0 + ((size_t)b * a.Cin + co) * HW) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int split = a.split > 0 ? a.split : a.Cout;
#pragma unroll 4
  for (int ci = grp; ci < a.Cout; ci += PW_GROUPS) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wt) + ci);
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int co = 0; co < CO; ++co)
      s.x = fmaf(w[co], g[co].x, s.x), s.y = fmaf(w[co], g[co].y, s.y), s.z = fmaf(w[co], g[co].z, s.z), s.w = fmaf(w[co], g[co].w, s.w);
    float* dst = ci < split ? a.out : a.out1;
    if (dst == nullptr) continue;
    const size_t o = ci < split ? ((size_t)b * split + ci) * HW : ((size_t)b * (a.Cout - split) + (ci - split)) * HW;
    reinterpret_cast<float4*>(dst + o)[q] = s;
  }
}

struct PwWgradArgs {
  VirtIn vin;
  const float* g;     // (B, Cout, HW)
  float* partial;     // [chunks][Cin][CO]
  int B, Cin, Cout, HW4, chunks_per_img;
};

constexpr int PW_CPW = 4;   // input channels per warp of the weight-gradient kernel (the gradient vectors are loaded once for all of them)

template <int CO>
__global__ void __launch_bounds__(PW_THREADS) pw_small_wgrad_kernel(const __grid_constant__ PwWgradArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci0 = (blockIdx.y * PW_GROUPS + warp) * PW_CPW;
  if (ci0 >= a.Cin) return;
  const int chunk = blockIdx.x, b = chunk / a.chunks_per_img, q0 = (chunk - b * a.chunks_per_img) * PW_CHUNK_Q;
  const int q1 = min(q0 + PW_CHUNK_Q, a.HW4);
  const size_t HW = (size_t)a.HW4 * 4;
  const float4* xp[PW_CPW];
#pragma unroll
  for (int j = 0; j < PW_CPW; ++j) xp[j] = reinterpret_cast<const float4*>(pw_plane(a.vin, b, min(ci0 + j, a.Cin - 1), HW));
  float s[PW_CPW][CO];
#pragma unroll
  for (int j = 0; j < PW_CPW; ++j)
#pragma unroll
    for (int co = 0; co < CO; ++co) s[j][co] = 0.f;
#pragma unroll 2
  for (int q = q0 + lane; q < q1; q += 32) {
    float4 v[PW_CPW], g[CO];
#pragma unroll
    for (int j = 0; j < PW_CPW; ++j) v[j] = __ldg(xp[j] + q);
#pragma unroll
    for (int co = 0; co < CO; ++co)
      g[co] = co < a.Cout ? __ldg(reinterpret_cast<const float4*>(a.g + ((size_t)b * a.Cout + co) * HW) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PW_CPW; ++j)
#pragma unroll
      for (int co = 0; co < CO; ++co) s[j][co] += (g[co].x * v[j].x + g[co].y * v[j].y) + (g[co].z * v[j].z + g[co].w * v[j].w);
  }
#pragma unroll
  for (int j = 0; j < PW_CPW; ++j)
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      const float t = warp_sum(s[j][co]);
      if (lane == 0 && ci0 + j < a.Cin) a.partial[((size_t)chunk * a.Cin + ci0 + j) * CO + co] = t;
    }
}

// gw[co][ci] = sum over chunks: one warp per (ci, co), lanes stride over the chunks, double accumulation in a fixed order
template <int CO>
__global__ void __launch_bounds__(128) pw_small_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int Cin, int Cout, float* __restrict__ gw) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // ci * CO + co
  if (i >= Cin * CO) return;
  const int ci = i / CO, co = i - ci * CO;
  if (co >= Cout) return;
  double s = 0.0;
  for (int k = lane; k < chunks; k += 32) s += (double)partial[(size_t)k * Cin * CO + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) gw[(size_t)co * Cin + ci] = (float)s;
}

static bool use_pw_small(int ks, int narrow_channels, int H, int W, int up0) {
  static const bool disabled = getenv("DD_NO_PW_SMALL") != nullptr;
  return !disabled && ks == 1 && narrow_channels <= 4 && ((size_t)H * W) % 4 == 0 && up0 == DD_UP_NONE;
}

static size_t pw_small_wgrad_bytes(int B, int H, int W, int Cin) {
  const int HW4 = (int)(((size_t)H * W) / 4);
  const int cpi = (HW4 + PW_CHUNK_Q - 1) / PW_CHUNK_Q;
  return (size_t)B * cpi * Cin * 4 * sizeof(float);
}

// weight table [Cin_f][4] of the layer (Cout_f <= 4 columns, zero padded) for both directions
static void pw_small_table(const float* w_oihw, float* wt_buf, int Cout_f, int Cin_f, cudaStream_t st) {
  const size_t wn = (size_t)Cin_f * 4;
  conv_prep_weights_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(w_oihw, wt_buf, Cout_f, Cin_f, 1, 4, 0);
  dd::count_launches(1);
}

// forward: args as prepared by conv_fwd_impl
static int run_pw_small_fwd(ConvArgs& a, float* wt_buf, const float* w_oihw, cudaStream_t st) {
  pw_small_table(w_oihw, wt_buf, a.Cout, a.Cin, st);
  a.wt = wt_buf;
  const int HW4 = (int)(((size_t)a.Ho * a.Wo) / 4);
  const bool split = a.Cin > 32;
  const dim3 grid(split ? (HW4 + 31) / 32 : (HW4 + PW_THREADS - 1) / PW_THREADS, 1, a.B);
#define DD_PW_FWD(CO)                                                                  \
  do {                                                                                 \
    if (split) pw_small_fwd_kernel<CO, true><<<grid, PW_THREADS, 0, st>>>(a);          \
    else pw_small_fwd_kernel<CO, false><<<grid, PW_THREADS, 0, st>>>(a);               \
  } while (0)
  if (a.Cout <= 1) DD_PW_FWD(1);
  else if (a.Cout <= 2) DD_PW_FWD(2);
  else DD_PW_FWD(4);
#undef DD_PW_FWD
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

// data gradient: a.Cin = the layer's output channels (<= 4), a.Cout = its input channels
static int run_pw_small_dgrad(ConvArgs& a, float* wt_buf, const float* w_oihw, cudaStream_t st) {
  pw_small_table(w_oihw, wt_buf, a.Cin, a.Cout, st);
  a.wt = wt_buf;
  const int HW4 = (int)(((size_t)a.Ho * a.Wo) / 4);
  const dim3 grid((HW4 + 31) / 32, 1, a.B);
  if (a.Cin <= 1) pw_small_dgrad_kernel<1><<<grid, PW_THREADS, 0, st>>>(a);
  else if (a.Cin <= 2) pw_small_dgrad_kernel<2><<<grid, PW_THREADS, 0, st>>>(a);
  else pw_small_dgrad_kernel<4><<<grid, PW_THREADS, 0, st>>>(a);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

static int run_pw_small_wgrad(const VirtIn& vin, const float* g, int B, int H, int W, int Cin, int Cout, float* partial, float* gw,
                              cudaStream_t st) {
  PwWgradArgs a;
  a.vin = vin, a.g = g, a.partial = partial, a.B = B, a.Cin = Cin, a.Cout = Cout;
  a.HW4 = (int)(((size_t)H * W) / 4);
  a.chunks_per_img = (a.HW4 + PW_CHUNK_Q - 1) / PW_CHUNK_Q;
  const int chunks = B * a.chunks_per_img;
  const dim3 grid(chunks, (Cin + PW_GROUPS * PW_CPW - 1) / (PW_GROUPS * PW_CPW));
  DD_REQUIRE(grid.y <= 65535, "pw_small_wgrad: too many input channels");
  const int nred = (Cin * 4 + 3) / 4;
  if (Cout <= 1) {
    pw_small_wgrad_kernel<1><<<grid, PW_THREADS, 0, st>>>(a);
    pw_small_wgrad_reduce_kernel<1><<<nred, 128, 0, st>>>(partial, chunks, Cin, Cout, gw);
  } else if (Cout <= 2) {
    pw_small_wgrad_kernel<2><<<grid, PW_THREADS, 0, st>>>(a);
    pw_small_wgrad_reduce_kernel<2><<<nred, 128, 0, st>>>(partial, chunks, Cin, Cout, gw);
  } else {
    pw_small_wgrad_kernel<4><<<grid, PW_THREADS, 0, st>>>(a);
    pw_small_wgrad_reduce_kernel<4><<<nred, 128, 0, st>>>(partial, chunks, Cin, Cout, gw);
  }
  dd::count_launches(2);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // namespace dd
