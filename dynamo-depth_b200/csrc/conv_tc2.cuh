// EXPERIMENTAL (DD_TC_CONV=2) -- written at the end of round 1 when the GPU budget was nearly spent.  Status: parity-green
// on the B200 (all 26 tests of tests/test_conv_gpu.py pass with DD_TC_CONV=2); one 3-repetition timing puts it level with
// the Winograd kernels on the >= 64-channel layers (0.77x .. 1.13x of their time) and behind on the 32-channel full-resolution
// layer (1.7x): producer-bound (8-channel K blocks, two groups), see DESIGN.md section 6.  Nothing selects it unless the
// environment variable is set.
//
// 3x3 decoder convolutions (forward and data gradient of ConvBlock / Conv3x3, networks/layers.py:85-121) as implicit GEMMs
// on the tcgen05 tensor cores (3xTF32, fp32 accuracy) WITHOUT the per-tap im2col gather of conv_tc.cuh, which keeps that
// kernel behind the Winograd path: the input patch of a tile is staged once per K block in shared memory and all nine taps
// become descriptor offsets into it.
//
//   output tile : 4 rows x 32 pixels of one image (M = 128: TMEM lane = 32 * row + pixel), BN = 32 or 64 output channels
//   K block     : 8 input channels
//   A           : the patch rows y-1 .. y+4, stored three times (dx = -1, 0, +1: source columns d .. d+31 at pixel
//                 positions 0 .. 31) as [copy d][patch row][channel][32 pixels] in the SWIZZLE_128B_BASE32B pattern
//                 (32-byte chunk ^= channel % 4), hi and lo: 2 x 18 KB.  Tap (dy, dx) = the MN-major descriptor at
//                 copy(dx+1) + (dy+1) * 1024 B with LBO = 1024 B (next output row = next patch row), SBO = 512 B.
//   B           : weights pre-split into hi / lo and pre-swizzled by conv_prep_tc2_weights_kernel: per (N tile, K block) one
//                 contiguous block [hi, lo][3][BN][32 k], k = 8 * (tap % 4) + channel within row tile tap / 4 (K-major
//                 SWIZZLE_128B rows; the K = 8 slice of an MMA is one tap) -- one cp.async.bulk per stage, no conversion.
//   per stage   : 9 taps x 3 splits = 27 MMAs (32 clocks each at N = 64); producers: 12 coalesced row loads and 72 scalar
//                 shared-memory stores per thread.
#pragma once
#include "tc_common.cuh"

namespace dd {

constexpr int C2_ROWS = 4, C2_COLS = 32;
constexpr int C2_CH = 8;
constexpr int C2_PROWS = C2_ROWS + 2;
constexpr int C2_ROW_BYTES = C2_CH * 128;                  // one patch row: [channel][32 pixels]
constexpr int C2_COPY_BYTES = C2_PROWS * C2_ROW_BYTES;     // 6144
constexpr int C2_A_BYTES = 3 * C2_COPY_BYTES;              // 18432 (hi or lo)
constexpr int C2_BJ = 3;                                   // weight row tiles of 32 k = 4 taps x 8 channels

struct ConvTc2Args {
  ConvArgs a;
  const float* wsplit;   // [n_tile][kb][hi, lo][3][BN][32], rows in SWIZZLE_128B order
  int kb_total;          // ceil(Cin / 8)
  int tiles_x, tiles_y, m_tiles, n_tiles;
  int stages;
};

__global__ void conv_prep_tc2_weights_kernel(const float* __restrict__ w, float* __restrict__ ws, int Cout_f, int Cin_f, int n_tiles,
                                             int kb_total, int BN, int transpose) {
  const int n_out = transpose ? Cin_f : Cout_f;
  const int n_in = transpose ? Cout_f : Cin_f;
  const size_t half = (size_t)C2_BJ * BN * 32;   // floats of one hi (or lo) block
  const size_t total = (size_t)n_tiles * kb_total * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % 32);
    const int n = (int)((i / 32) % BN);
    const int j = (int)((i / (32 * (size_t)BN)) % C2_BJ);
    const size_t blk = i / half;                 // nt * kb_total + kb
    const int kb = (int)(blk % kb_total), nt = (int)(blk / kb_total);
    const int kk = (((e >> 2) ^ (n & 7)) << 2) | (e & 3);   // physical position e of the row holds logical k = kk
    const int tap = 4 * j + (kk >> 3), ci = kb * C2_CH + (kk & 7), co = nt * BN + n;
    float v = 0.f;
    if (tap < 9 && ci < n_in && co < n_out)
      v = transpose ? __ldg(w + ((size_t)ci * Cin_f + co) * 9 + (8 - tap)) : __ldg(w + ((size_t)co * Cin_f + ci) * 9 + tap);
    const float hi = tc::tf32_rn(v);
    float* dst = ws + blk * 2 * half + (size_t)(j * BN + n) * 32 + e;
    dst[0] = hi;
    dst[half] = v - hi;
  }
}

template <int NB32>
__global__ void __launch_bounds__(tc::cta_threads(2), 1) conv_tc2_kernel(const __grid_constant__ ConvTc2Args g) {
  using namespace tc;
  constexpr int BN = NB32 * 32;
  constexpr int B_BYTES = C2_BJ * BN * 128;                 // hi or lo
  constexpr int STAGE_BYTES = 2 * C2_A_BYTES + 2 * B_BYTES; // [A_hi][A_lo][B_hi][B_lo]
  constexpr int G = 2;
  constexpr int EPI_WARP0 = epi_warp0(G), MMA_WARP = mma_warp(G);
  const ConvArgs& a = g.a;

  extern __shared__ uint8_t smem_raw[];
  // `full` = 128 producer arrivals + the arrive.expect_tx that announces the bulk copy of the weights
  const Cta c = cta_setup(smem_raw, g.stages, STAGE_BYTES, MMA_WARP, GROUP_THREADS + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;
  const int tiles_img = g.tiles_x * g.tiles_y;

  if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ producers
    const int grp = warp >> 2, ptid = threadIdx.x & (GROUP_THREADS - 1), pw = ptid >> 5;
    const uint32_t stages = (uint32_t)c.stages;
    const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
    const int C0 = a.vin.C0, Call = a.vin.C0 + a.vin.C1;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
      const int b = mt / tiles_img, rem = mt - b * tiles_img, ty = rem / g.tiles_x;
      const int y0 = ty * C2_ROWS, x0 = (rem - ty * g.tiles_x) * C2_COLS;
      // source offsets (inside one channel plane; -1 = padding / outside) of patch column `lane` and, for lanes 0 and 1,
      // of the two extra columns 32 + lane, for the six patch rows
      int o0[C2_PROWS], o1[C2_PROWS], e0[C2_PROWS], e1[C2_PROWS];
#pragma unroll
      for (int r = 0; r < C2_PROWS; ++r) {
        TapEntry te;
        build_tile_map(a.vin, a.oy + y0 - 1 + r, a.ox + x0 - 1 + lane, te, o1[r]);
        o0[r] = te.o00;
        e0[r] = e1[r] = -1;
        if (lane < 2) {
          build_tile_map(a.vin, a.oy + y0 - 1 + r, a.ox + x0 - 1 + 32 + lane, te, e1[r]);
          e0[r] = te.o00;
        }
      }
      const float* img0 = a.vin.x0 + (size_t)b * C0 * plane0;
      const float* img1 = a.vin.x1 + (size_t)b * a.vin.C1 * plane1;   // only dereferenced when C1 > 0
      for (int kb = 0; kb < g.kb_total; ++kb, ++it) {
        if ((int)(it % G) != grp) continue;
        const uint32_t stage = it % stages, ph = (it / stages) & 1u;
        // 12 (patch row, channel) pairs per thread: channels 2 pw and 2 pw + 1 of the block, all six rows
        float v[12], ve[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int r = i % C2_PROWS, ch = kb * C2_CH + 2 * pw + i / C2_PROWS;
          v[i] = ve[i] = 0.f;
          if (ch < C0) {
            if (o0[r] >= 0) v[i] = __ldg(img0 + (size_t)ch * plane0 + o0[r]);
            if (e0[r] >= 0) ve[i] = __ldg(img0 + (size_t)ch * plane0 + e0[r]);
          } else if (ch < Call) {
            if (o1[r] >= 0) v[i] = __ldg(img1 + (size_t)(ch - C0) * plane1 + o1[r]);
            if (e1[r] >= 0) ve[i] = __ldg(img1 + (size_t)(ch - C0) * plane1 + e1[r]);
          }
        }
        mbar_wait(c.empty_bar + 8 * stage, ph ^ 1u);
        const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, b_hi = a_hi + 2 * C2_A_BYTES;
        if (ptid == 0) {   // weights of this (N tile, K block): one bulk copy, completion counted on `full`
          const float* src = g.wsplit + ((size_t)nt * g.kb_total + kb) * (2 * B_BYTES / 4);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c.full_bar + 8 * stage), "r"(2 * B_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(b_hi), "l"(src),
                       "r"(2 * B_BYTES), "r"(c.full_bar + 8 * stage)
                       : "memory");
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int r = i % C2_PROWS, chl = 2 * pw + i / C2_PROWS;   // channel inside the block
          const uint32_t row = a_hi + (uint32_t)(r * C2_ROW_BYTES + chl * 128);
          const float hi = tf32_rn(v[i]), lo = v[i] - hi;
#pragma unroll
          for (int d = 0; d < 3; ++d) {   // source column `lane` is pixel lane - d of copy d
            const int p = lane - d;
            if (p >= 0) {
              const uint32_t o = row + (uint32_t)(d * C2_COPY_BYTES) + (uint32_t)((((p >> 3) ^ (chl & 3)) << 5) | ((p & 7) << 2));
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(o), "f"(hi) : "memory");
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(o + C2_A_BYTES), "f"(lo) : "memory");
            }
          }
          if (lane < 2) {   // source columns 32 and 33: pixels 32 + lane - d of the copies d > lane
            const float hie = tf32_rn(ve[i]), loe = ve[i] - hie;
#pragma unroll
            for (int d = 1; d < 3; ++d) {
              const int p = 32 + lane - d;
              if (p < 32) {
                const uint32_t o = row + (uint32_t)(d * C2_COPY_BYTES) + (uint32_t)((((p >> 3) ^ (chl & 3)) << 5) | ((p & 7) << 2));
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(o), "f"(hie) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(o + C2_A_BYTES), "f"(loe) : "memory");
              }
            }
          }
        }
        fence_async_smem();
        mbar_arrive(c.full_bar + 8 * stage);
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);   // D fp32, A / B tf32, A MN-major, B K-major
      const uint32_t stages = (uint32_t)c.stages;
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
        mbar_wait(c.tempty_bar + 8 * buf, tph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = c.tmem_base + buf * 256u;
        for (int kb = 0; kb < g.kb_total; ++kb, ++it) {
          const uint32_t stage = it % stages, ph = (it / stages) & 1u;
          mbar_wait(c.full_bar + 8 * stage, ph);
          tc_fence_after();
          const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, a_lo = a_hi + C2_A_BYTES;
          const uint32_t b_hi = a_lo + C2_A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int t = 0; t < 9; ++t) {   // tap (ky, kx) = (t / 3, t % 3): patch rows ky .. ky + 3 of copy kx
            const uint32_t oa = (uint32_t)((t % 3) * C2_COPY_BYTES + (t / 3) * C2_ROW_BYTES);
            const uint32_t ob = (uint32_t)((t >> 2) * (BN * 128) + (t & 3) * 32);
            const uint64_t da_hi = umma_desc<true>(a_hi + oa, C2_ROW_BYTES), da_lo = umma_desc<true>(a_lo + oa, C2_ROW_BYTES);
            const uint64_t db_hi = umma_desc<false>(b_hi + ob), db_lo = umma_desc<false>(b_lo + ob);
            umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > 0 || t > 0) ? 1u : 0u);
            umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
            umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
          }
          tc_commit(c.empty_bar + 8 * stage);
        }
        tc_commit(c.tfull_bar + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM lane = 32 * tile row + pixel
    const int ew = warp - EPI_WARP0;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
      const int b = mt / tiles_img, rem = mt - b * tiles_img, ty = rem / g.tiles_x;
      const int y = ty * C2_ROWS + ew, x = (rem - ty * g.tiles_x) * C2_COLS + lane;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(c.tfull_bar + 8 * buf, tph);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < NB32; ++cb) {
        const int co0 = nt * BN + cb * 32;
        if (co0 >= a.Cout) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld32(c.tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 256u + (uint32_t)(cb * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) emit_output(a, b, co0 + j, y, x, __uint_as_float(r[j]));   // drops y >= Ho, x >= Wo, co >= Cout
      }
      tc_fence_before();
      mbar_arrive(c.tempty_bar + 8 * buf);
    }
  }
  cta_teardown(c, MMA_WARP);
}

template <int NB32>
static int launch_conv_tc2(ConvTc2Args& g, int sms, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * C2_A_BYTES + 2 * C2_BJ * NB32 * 32 * 128;
  g.stages = tc::stages_for(STAGE_BYTES);
  const int smem = tc::smem_bytes(g.stages, STAGE_BYTES);
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<NB32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BUDGET));
    configured = true;
  }
  const int total = g.m_tiles * g.n_tiles;
  conv_tc2_kernel<NB32><<<total < sms ? total : sms, tc::cta_threads(2), smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(g);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

static bool use_tc2_conv(int ks, int cin, int cout) {
  static const char* env = getenv("DD_TC_CONV");
  return env != nullptr && env[0] == '2' && ks == 3 && cout > 16 && cin >= 8;
}

// bytes of prepared weights run_conv_tc2 writes (conv_ws sizes the workspace with it)
static size_t conv_tc2_weight_bytes(int cin, int cout) {
  const int BN = cout <= 32 ? 32 : 64;
  return (size_t)((cout + BN - 1) / BN) * ((cin + C2_CH - 1) / C2_CH) * 2 * C2_BJ * BN * 32 * sizeof(float);
}

static int run_conv_tc2(const ConvArgs& args, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose, int sms,
                        cudaStream_t st) {
  DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_tc2_kernel: bilinear up-sampling must be materialised first");
  ConvTc2Args g;
  memset(&g, 0, sizeof(g));
  g.a = args;
  const int BN = args.Cout <= 32 ? 32 : 64;
  g.n_tiles = (args.Cout + BN - 1) / BN;
  g.kb_total = (args.Cin + C2_CH - 1) / C2_CH;
  const size_t wn = (size_t)g.n_tiles * g.kb_total * C2_BJ * BN * 32;
  conv_prep_tc2_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(w_oihw, wt_buf, Cout_f, Cin_f, g.n_tiles,
                                                                                                     g.kb_total, BN, transpose ? 1 : 0);
  dd::count_launches(1);
  g.wsplit = wt_buf;
  g.tiles_x = (args.Wo + C2_COLS - 1) / C2_COLS;
  g.tiles_y = (args.Ho + C2_ROWS - 1) / C2_ROWS;
  g.m_tiles = args.B * g.tiles_x * g.tiles_y;
  return BN == 32 ? launch_conv_tc2<1>(g, sms, st) : launch_conv_tc2<2>(g, sms, st);
}

}  // namespace dd
