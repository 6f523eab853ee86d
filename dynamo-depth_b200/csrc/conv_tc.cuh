// Decoder convolutions (3x3 and 1x1 with more than 16 output channels: forward and data gradient of ConvBlock /
// Conv3x3, networks/layers.py:85-121, as used by depth_decoder.py:40-55,99-115, motion_decoder.py:24-62,
// pose_decoder.py:16-37) as implicit GEMMs on the tcgen05 tensor cores at fp32 accuracy (3xTF32 operand split, fp32
// accumulation in TMEM; building blocks in tc_common.cuh).  Included by conv.cu after ConvArgs / build_tile_map /
// emit_output.
//
//   D (128 output pixels x BN output channels) += A (pixels x 32 reduction steps) . B (channels x 32 steps)^T
//   reduction index k = tap * cin_pad + ci  (tap-major, Cin padded to 32): one K block = one filter tap, 32 channels
//
// A is MN-major (the pixel index is the contiguous one, as in NCHW memory): a producer thread owns 4 consecutive
// output pixels of the tile (linear index over batch x Ho x Wo, so every level fills its tiles completely), derives
// their source offsets for the block's tap once (reflect / zero padding, nearest up-sampling, skip concatenation via
// build_tile_map) and then only adds one channel plane per element -- 8 channels x 4 pixels per thread and K block.
// B = prepared weights [Cout][tap][cin_pad] (K-major rows), plain 16-byte loads.  The epilogue reads TMEM lane = pixel,
// so for every output channel a warp writes 32 consecutive pixels of one NCHW plane (coalesced without a transpose) through
// emit_output (bias, ELU / sigmoid / ReLU, residual, split destinations of the data gradient).
#pragma once
#include "tc_common.cuh"

namespace dd {

struct ConvTcArgs {
  ConvArgs a;        // a.wt = prepared weights [Cout][KK][cin_pad]
  int cin_pad;       // Cin rounded up to a multiple of 32
  int kb_per_tap;    // cin_pad / 32
  int kb_total;      // KK * kb_per_tap
  int m_tiles, n_tiles;
  int P;             // B * Ho * Wo output pixels
  int stages;
};

// weights OIHW (Cout_f, Cin_f, k, k) -> wt[co][tap][cin_pad]; transpose: the data-gradient convolution (input and output
// channels swapped, taps flipped)
__global__ void conv_prep_tc_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout_f, int Cin_f, int KK,
                                            int cin_pad, int transpose) {
  const int n_out = transpose ? Cin_f : Cout_f;   // rows of wt
  const int n_in = transpose ? Cout_f : Cin_f;    // valid reduction channels
  const size_t total = (size_t)n_out * KK * cin_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int tap = (int)((i / cin_pad) % KK);
    const int co = (int)(i / ((size_t)cin_pad * KK));
    float v = 0.f;
    if (ci < n_in) {
      if (!transpose) v = __ldg(w + ((size_t)co * Cin_f + ci) * KK + tap);
      else v = __ldg(w + ((size_t)ci * Cin_f + co) * KK + (KK - 1 - tap));
    }
    wt[i] = v;
  }
}

template <int NB32>
__host__ __device__ constexpr int conv_tc_groups() { return tc::groups_for(tc::stages_for(2 * tc::A_TILE_BYTES + 2 * NB32 * 32 * tc::BK * 4)); }

template <int KS, int NB32>
__global__ void __launch_bounds__(tc::cta_threads(conv_tc_groups<NB32>()), 1) conv_tc_kernel(const __grid_constant__ ConvTcArgs g) {
  using namespace tc;
  constexpr int BN = NB32 * 32;
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  constexpr int G = conv_tc_groups<NB32>();
  constexpr int EPI_WARP0 = epi_warp0(G), MMA_WARP = mma_warp(G);
  constexpr int A_F4 = BM * BK / 4 / GROUP_THREADS;   // 8
  constexpr int B_F4 = BN * BK / 4 / GROUP_THREADS;   // 2 * NB32
  const ConvArgs& a = g.a;

  extern __shared__ uint8_t smem_raw[];
  const Cta c = cta_setup(smem_raw, g.stages, STAGE_BYTES, MMA_WARP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;
  const int HoWo = a.Ho * a.Wo;

  if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ producers
    const int grp = warp >> 2, ptid = threadIdx.x & (GROUP_THREADS - 1);
    const int c4 = ptid & 31, kq = ptid >> 5;   // pixels 4*c4 .. 4*c4+3 of the tile; reduction steps kq, kq+4, ... (TileMap<true, 128>)
    TileMap<true, BM> ma;
    TileMap<false, BN> mb;
    ma.init(ptid), mb.init(ptid);
    const uint32_t stages = (uint32_t)c.stages, groups = G;
    const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
    const int C0 = a.vin.C0, Call = a.vin.C0 + a.vin.C1;
    const long long Ktot = (long long)KS * KS * g.cin_pad;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
      int py[4], px[4];
      const float* base0[4];
      const float* base1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = mt * BM + c4 * 4 + j;
        const int b = p / HoWo, rem = p - b * HoWo, y = rem / a.Wo;
        py[j] = p < g.P ? y : -(1 << 20);   // rows past the end read zeros
        px[j] = rem - y * a.Wo;
        base0[j] = a.vin.x0 + (size_t)b * C0 * plane0;
        base1[j] = a.vin.x1 + (size_t)b * a.vin.C1 * plane1;   // only dereferenced when C1 > 0
      }
      int last_tap = -1;
      const float* p0[4];
      const float* p1[4];
      for (int kb = 0; kb < g.kb_total; ++kb, ++it) {
        if ((int)(it % groups) != grp) continue;
        const uint32_t stage = it % stages, ph = (it / stages) & 1u;
        const int tap = kb / g.kb_per_tap, cb0 = (kb - tap * g.kb_per_tap) * BK;
        if (tap != last_tap) {   // source positions of the four pixels for this tap (NULL = padding / outside)
          last_tap = tap;
          const int dy = KS == 3 ? tap / 3 - 1 : 0, dx = KS == 3 ? tap - (tap / 3) * 3 - 1 : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            TapEntry e;
            int o1;
            build_tile_map(a.vin, a.oy + py[j] + dy, a.ox + px[j] + dx, e, o1);
            p0[j] = e.o00 >= 0 ? base0[j] + e.o00 : nullptr;
            p1[j] = (o1 >= 0 && a.vin.C1 > 0) ? base1[j] + o1 : nullptr;
          }
        }
        float4 va[A_F4], vb[B_F4];
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
          const int ch = cb0 + kq + 4 * i;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (ch < C0) {
            const size_t co = (size_t)ch * plane0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (p0[j]) v[j] = __ldg(p0[j] + co);
          } else if (ch < Call) {
            const size_t co = (size_t)(ch - C0) * plane1;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (p1[j]) v[j] = __ldg(p1[j] + co);
          }
          va[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
        load_tile<false, BN, B_F4>(vb, mb, a.wt, Ktot, nt * BN, a.Cout, kb * BK, (int)Ktot);
        mbar_wait(c.empty_bar + 8 * stage, ph ^ 1u);
        const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
        const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
        store_tile<true, BM, A_F4>(va, ma, a_hi, a_lo);    // chunk i = (step kq + 4 i, pixels 4 c4 ..)
        store_tile<false, BN, B_F4>(vb, mb, b_hi, b_lo);
        fence_async_smem();
        mbar_arrive(c.full_bar + 8 * stage);
      }
    }
  } else if (warp == MMA_WARP) {
    mma_issue_loop<true, false, BN>(c, total_tiles, total_tiles, g.kb_total, g.kb_total);
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM lane = pixel
    const int ew = warp - EPI_WARP0;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      const int p = mt * BM + ew * 32 + lane;
      const int b = p / HoWo, rem = p - b * HoWo;
      const int y = p < g.P ? rem / a.Wo : a.Ho;   // emit_output drops y >= Ho
      const int x = rem - (rem / a.Wo) * a.Wo;
      mbar_wait(c.tfull_bar + 8 * buf, tph);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < NB32; ++cb) {
        const int co0 = nt * BN + cb * 32;
        if (co0 >= a.Cout) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld32(c.tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 256u + (uint32_t)(cb * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) emit_output(a, b, co0 + j, y, x, __uint_as_float(r[j]));
      }
      tc_fence_before();
      mbar_arrive(c.tempty_bar + 8 * buf);
    }
  }
  cta_teardown(c, MMA_WARP);
}

template <int KS, int NB32>
static int launch_conv_tc(ConvTcArgs& g, int sms, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * tc::A_TILE_BYTES + 2 * NB32 * 32 * tc::BK * 4;
  g.stages = tc::stages_for(STAGE_BYTES);
  const int smem = tc::smem_bytes(g.stages, STAGE_BYTES);
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KS, NB32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BUDGET));
    configured = true;
  }
  const int total = g.m_tiles * g.n_tiles;
  // > half of the SM's shared memory in every configuration: one CTA per SM owns all 512 TMEM columns
  conv_tc_kernel<KS, NB32><<<total < sms ? total : sms, tc::cta_threads(conv_tc_groups<NB32>()), smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(g);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

template <int KS>
static int launch_conv_tc_nb(ConvTcArgs& g, int nb32, int sms, cudaStream_t st) {
  switch (nb32) {
    case 1: return launch_conv_tc<KS, 1>(g, sms, st);
    case 2: return launch_conv_tc<KS, 2>(g, sms, st);
    case 3: return launch_conv_tc<KS, 3>(g, sms, st);
    case 4: return launch_conv_tc<KS, 4>(g, sms, st);
    case 5: return launch_conv_tc<KS, 5>(g, sms, st);
    case 6: return launch_conv_tc<KS, 6>(g, sms, st);
    case 7: return launch_conv_tc<KS, 7>(g, sms, st);
    default: return launch_conv_tc<KS, 8>(g, sms, st);
  }
}

// Opt-in (DD_TC_CONV=1): parity-green, but the scalar im2col gather of the producers keeps it behind the Winograd
// kernels (B200: 1.45 ms vs 1.05 ms on the motion level-4 layer), so the default stays on the CUDA-core path.
static bool use_tc_conv(int ks, int cin, int cout) {
  static const bool enabled = getenv("DD_TC_CONV") != nullptr;
  return enabled && (ks == 3 || ks == 1) && cout > 16 && cin >= 8;
}

// args: virtual input / output grid / epilogue filled in; prepares the weights into wt_buf and launches
static int run_conv_tc(const ConvArgs& args, int ks, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose,
                       int sms, cudaStream_t st) {
  DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_tc_kernel: bilinear up-sampling must be materialised first");
  ConvTcArgs g;
  memset(&g, 0, sizeof(g));
  g.a = args;
  const int KK = ks * ks;
  g.cin_pad = (args.Cin + 31) / 32 * 32;
  g.kb_per_tap = g.cin_pad / 32;
  g.kb_total = KK * g.kb_per_tap;
  const size_t wn = (size_t)args.Cout * KK * g.cin_pad;
  conv_prep_tc_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(w_oihw, wt_buf, Cout_f, Cin_f, KK,
                                                                                                    g.cin_pad, transpose ? 1 : 0);
  dd::count_launches(1);
  g.a.wt = wt_buf;
  const long long P = (long long)args.B * args.Ho * args.Wo;
  DD_REQUIRE(P < (1ll << 30), "conv_tc_kernel: too many output pixels (%lld)", P);
  g.P = (int)P;
  g.m_tiles = (g.P + tc::BM - 1) / tc::BM;
  const int n32 = (args.Cout + 31) / 32;
  g.n_tiles = (n32 + 7) / 8;
  const int nb32 = (n32 + g.n_tiles - 1) / g.n_tiles;
  g.n_tiles = (n32 + nb32 - 1) / nb32;
  return ks == 3 ? launch_conv_tc_nb<3>(g, nb32, sms, st) : launch_conv_tc_nb<1>(g, nb32, sms, st);
}

}  // namespace dd
