// EXPERIMENTAL (DD_TC_CONV=3) -- written after round 1's GPU budget was spent: compiles for sm_100a, has NOT run on hardware.
// First thing to do with it: `DD_TC_CONV=3 python -m pytest tests/test_conv_gpu.py -m gpu`, then
// `DD_TC_CONV=3 python dev/kernel_bench.py --what conv` against DD_TC_CONV=2 and the Winograd default.
//
// Weight-stationary variant of conv_tc2.cuh (same tile geometry, same operand layouts, same prepared weights): conv_tc2
// re-reads the 49 KB hi / lo weight block of every K block for every 128-pixel tile (24 FLOP per L2 byte).  Here a CTA works
// on FOUR pixel tiles at once: the weights of a K block are loaded once (their own 2-stage ring) and the 27 MMAs run for
// each of the four tiles, whose accumulators sit side by side in TMEM (4 x BN columns; two such sets = 512 columns, so the
// epilogue of one set overlaps the MMAs of the next).  Only the 36 KB patches stream through a 3-stage ring:
//   shared memory = 2 x 49 KB (weights) + 3 x 36 KB (patches) = 206 KB;  weight traffic per FLOP: 1/4 of conv_tc2.
#pragma once
#include "conv_tc2.cuh"

namespace dd {

constexpr int C3_TILES = 4;        // pixel tiles per work item (accumulators side by side in TMEM)
constexpr int C3_WSTAGES = 2;
constexpr int C3_PSTAGES = 3;
constexpr int C3_PSTAGE_BYTES = 2 * C2_A_BYTES;   // patch hi + lo

struct ConvTc3Args {
  ConvArgs a;
  const float* wsplit;   // as conv_tc2: [n_tile][kb][hi, lo][3][BN][32]
  int kb_total;
  int tiles_x, tiles_y, m_tiles, n_tiles;
  int super_tiles;       // ceil(m_tiles / 4)
};

template <int NB32>
__global__ void __launch_bounds__(tc::cta_threads(2), 1) conv_tc3_kernel(const __grid_constant__ ConvTc3Args g) {
  using namespace tc;
  constexpr int BN = NB32 * 32;
  constexpr int B_BYTES = C2_BJ * BN * 128;          // hi or lo
  constexpr int WSTAGE_BYTES = 2 * B_BYTES;
  constexpr int G = 2;
  constexpr int EPI_WARP0 = epi_warp0(G), MMA_WARP = mma_warp(G);
  const ConvArgs& a = g.a;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const uint32_t w_base = smem_base, p_base = smem_base + C3_WSTAGES * WSTAGE_BYTES;
  const uint32_t bar_base = p_base + C3_PSTAGES * C3_PSTAGE_BYTES;
  const uint32_t wfull = bar_base, wempty = bar_base + 16, pfull = bar_base + 32, pempty = bar_base + 56;
  const uint32_t tfull = bar_base + 80, tempty = bar_base + 96;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C3_WSTAGES * WSTAGE_BYTES + C3_PSTAGES * C3_PSTAGE_BYTES + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C3_WSTAGES; ++s) mbar_init(wfull + 8 * s, 1), mbar_init(wempty + 8 * s, 1);
    for (int s = 0; s < C3_PSTAGES; ++s) mbar_init(pfull + 8 * s, GROUP_THREADS), mbar_init(pempty + 8 * s, 1);
    for (int s = 0; s < 2; ++s) mbar_init(tfull + 8 * s, 1), mbar_init(tempty + 8 * s, 4 * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = g.super_tiles * g.n_tiles;
  const int tiles_img = g.tiles_x * g.tiles_y;

  if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ producers (group grp fills tiles grp and grp + 2)
    const int grp = warp >> 2, ptid = threadIdx.x & (GROUP_THREADS - 1), pw = ptid >> 5;
    const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
    const int C0 = a.vin.C0, Call = a.vin.C0 + a.vin.C1;
    uint32_t pit = 0, wit = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int nt = work % g.n_tiles, st = work / g.n_tiles;
      int o0[2][C2_PROWS], o1[2][C2_PROWS], e0[2][C2_PROWS], e1[2][C2_PROWS];
      const float* img0[2];
      const float* img1[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int mt = st * C3_TILES + grp + 2 * s;
        const bool valid = mt < g.m_tiles;           // tiles past the end are filled with zeros and never stored
        const int b = valid ? mt / tiles_img : 0, rem = valid ? mt - b * tiles_img : 0, ty = rem / g.tiles_x;
        const int y0 = valid ? ty * C2_ROWS : (1 << 20), x0 = (rem - ty * g.tiles_x) * C2_COLS;
        img0[s] = a.vin.x0 + (size_t)b * C0 * plane0;
        img1[s] = a.vin.x1 + (size_t)b * a.vin.C1 * plane1;
#pragma unroll
        for (int r = 0; r < C2_PROWS; ++r) {
          TapEntry te;
          build_tile_map(a.vin, a.oy + y0 - 1 + r, a.ox + x0 - 1 + lane, te, o1[s][r]);
          o0[s][r] = te.o00;
          e0[s][r] = e1[s][r] = -1;
          if (lane < 2) {
            build_tile_map(a.vin, a.oy + y0 - 1 + r, a.ox + x0 - 1 + 32 + lane, te, e1[s][r]);
            e0[s][r] = te.o00;
          }
        }
      }
      for (int kb = 0; kb < g.kb_total; ++kb, ++wit) {
#pragma unroll
        for (int t = 0; t < C3_TILES; ++t, ++pit) {
          if ((t & 1) != grp) continue;
          const int s = t >> 1;   // compile-time after unrolling: which of the group's two tiles
          const uint32_t ps = pit % C3_PSTAGES, pph = (pit / C3_PSTAGES) & 1u;
          float v[12], ve[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int r = i % C2_PROWS, ch = kb * C2_CH + 2 * pw + i / C2_PROWS;
            v[i] = ve[i] = 0.f;
            if (ch < C0) {
              if (o0[s][r] >= 0) v[i] = __ldg(img0[s] + (size_t)ch * plane0 + o0[s][r]);
              if (e0[s][r] >= 0) ve[i] = __ldg(img0[s] + (size_t)ch * plane0 + e0[s][r]);
            } else if (ch < Call) {
              if (o1[s][r] >= 0) v[i] = __ldg(img1[s] + (size_t)(ch - C0) * plane1 + o1[s][r]);
              if (e1[s][r] >= 0) ve[i] = __ldg(img1[s] + (size_t)(ch - C0) * plane1 + e1[s][r]);
            }
          }
          if (t == 0 && ptid == 0) {   // (group 0 only) weights of this K block: one bulk copy into the weight ring
            const uint32_t ws = wit % C3_WSTAGES, wph = (wit / C3_WSTAGES) & 1u;
            mbar_wait(wempty + 8 * ws, wph ^ 1u);
            const float* src = g.wsplit + ((size_t)nt * g.kb_total + kb) * (WSTAGE_BYTES / 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wfull + 8 * ws), "r"(WSTAGE_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_base + ws * WSTAGE_BYTES),
                         "l"(src), "r"(WSTAGE_BYTES), "r"(wfull + 8 * ws)
                         : "memory");
          }
          mbar_wait(pempty + 8 * ps, pph ^ 1u);
          const uint32_t a_hi = p_base + ps * C3_PSTAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int r = i % C2_PROWS, chl = 2 * pw + i / C2_PROWS;
            const uint32_t row = a_hi + (uint32_t)(r * C2_ROW_BYTES + chl * 128);
            const float hi = tf32_rn(v[i]), lo = v[i] - hi;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const int p = lane - d;
              if (p >= 0) {
                const uint32_t o = row + (uint32_t)(d * C2_COPY_BYTES) + (uint32_t)((((p >> 3) ^ (chl & 3)) << 5) | ((p & 7) << 2));
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(o), "f"(hi) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(o + C2_A_BYTES), "f"(lo) : "memory");
              }
            }
            if (lane < 2) {
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
          mbar_arrive(pfull + 8 * ps);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      uint32_t pit = 0, wit = 0, tcount = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
        mbar_wait(tempty + 8 * buf, tph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < g.kb_total; ++kb, ++wit) {
          const uint32_t ws = wit % C3_WSTAGES, wph = (wit / C3_WSTAGES) & 1u;
          mbar_wait(wfull + 8 * ws, wph);
          const uint32_t b_hi = w_base + ws * WSTAGE_BYTES, b_lo = b_hi + B_BYTES;
          for (int t = 0; t < C3_TILES; ++t, ++pit) {
            const uint32_t ps = pit % C3_PSTAGES, pph = (pit / C3_PSTAGES) & 1u;
            mbar_wait(pfull + 8 * ps, pph);
            tc_fence_after();
            const uint32_t a_hi = p_base + ps * C3_PSTAGE_BYTES, a_lo = a_hi + C2_A_BYTES;
            const uint32_t d_tmem = tmem_base + buf * 256u + (uint32_t)(t * BN);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t oa = (uint32_t)((tap % 3) * C2_COPY_BYTES + (tap / 3) * C2_ROW_BYTES);
              const uint32_t ob = (uint32_t)((tap >> 2) * (BN * 128) + (tap & 3) * 32);
              const uint64_t da_hi = umma_desc<true>(a_hi + oa, C2_ROW_BYTES), da_lo = umma_desc<true>(a_lo + oa, C2_ROW_BYTES);
              const uint64_t db_hi = umma_desc<false>(b_hi + ob), db_lo = umma_desc<false>(b_lo + ob);
              umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > 0 || tap > 0) ? 1u : 0u);
              umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
            }
            tc_commit(pempty + 8 * ps);
          }
          tc_commit(wempty + 8 * ws);   // all four tiles have consumed the weights of this K block
        }
        tc_commit(tfull + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: four accumulators per work item
    const int ew = warp - EPI_WARP0;
    uint32_t tcount = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++tcount) {
      const int nt = work % g.n_tiles, st = work / g.n_tiles;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(tfull + 8 * buf, tph);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < C3_TILES; ++t) {
        const int mt = st * C3_TILES + t;
        if (mt >= g.m_tiles) break;   // warp-uniform
        const int b = mt / tiles_img, rem = mt - b * tiles_img, ty = rem / g.tiles_x;
        const int y = ty * C2_ROWS + ew, x = (rem - ty * g.tiles_x) * C2_COLS + lane;
#pragma unroll 1
        for (int cb = 0; cb < NB32; ++cb) {
          const int co0 = nt * BN + cb * 32;
          if (co0 >= a.Cout) break;   // warp-uniform
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 256u + (uint32_t)(t * BN + cb * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) emit_output(a, b, co0 + j, y, x, __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int NB32>
static int launch_conv_tc3(ConvTc3Args& g, int sms, cudaStream_t st) {
  constexpr int SMEM = 1024 + C3_WSTAGES * 2 * C2_BJ * NB32 * 32 * 128 + C3_PSTAGES * C3_PSTAGE_BYTES + 256;
  static_assert(SMEM <= tc::SMEM_BUDGET, "conv_tc3: shared memory budget");
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<NB32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BUDGET));
    configured = true;
  }
  const int total = g.super_tiles * g.n_tiles;
  // > half of the SM's shared memory in every configuration: one CTA per SM owns all 512 TMEM columns
  conv_tc3_kernel<NB32><<<total < sms ? total : sms, tc::cta_threads(2), SMEM < 120 * 1024 ? 120 * 1024 : SMEM, st>>>(g);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

static bool use_tc3_conv(int ks, int cin, int cout) {
  static const char* env = getenv("DD_TC_CONV");
  return env != nullptr && env[0] == '3' && ks == 3 && cout > 16 && cin >= 8;
}

static int run_conv_tc3(const ConvArgs& args, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose, int sms,
                        cudaStream_t st) {
  DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_tc3_kernel: bilinear up-sampling must be materialised first");
  ConvTc3Args g;
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
  g.super_tiles = (g.m_tiles + C3_TILES - 1) / C3_TILES;
  return BN == 32 ? launch_conv_tc3<1>(g, sms, st) : launch_conv_tc3<2>(g, sms, st);
}

}  // namespace dd
