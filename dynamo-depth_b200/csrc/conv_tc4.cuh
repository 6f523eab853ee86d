// 3x3 decoder convolutions (forward and data gradient of ConvBlock / Conv3x3, networks/layers.py:85-121 in the reference, as used
// by depth_decoder.py:40-55,99-115, motion_decoder.py:34-62, pose_decoder.py:16-37) as implicit GEMMs on the tcgen05 tensor
// cores: 3xTF32 (fp32 accuracy), accumulators in TMEM, weight-stationary over several pixel tiles.
//
// Formulation ("linear patch").  A CTA stages, per K block of 8 input channels, a patch of PR x PW positions of the VIRTUAL
// input (padding, nearest up-sampling and skip concatenation resolved by the loader) in shared memory as
//
//      [4-channel chunk][linear position p = r * PW + c][4 channels]          (16 bytes per position and chunk)
//
// which is exactly the canonical K-major NO-SWIZZLE operand layout of tcgen05.mma (rows 16 bytes apart, 8-row groups
// contiguous = SBO 128 B, the two 16-byte K chunks of a tf32 MMA LBO bytes apart).  The output at linear position p is
//      sum over taps (dy, dx) of  W[tap] . X[p + dy * PW + dx]
// so the A operand of tap (dy, dx) for the M tile of 128 consecutive positions starting at p0 is THE SAME patch addressed
// at a start address (p0 + dy * PW + dx) * 16 bytes further on: all nine taps are descriptor offsets, every input element
// is split into hi / lo and written to shared memory exactly once (two 16-byte stores per position and chunk), and no
// dx-shifted copies or swizzle arithmetic are needed.  The price: columns 0 and PW-1 of every patch row are halo, i.e. 2 of
// PW accumulator rows are computed and dropped (3 % at PW = 64).
//
//   M tile     : 128 consecutive linear positions (PW = 64: two output rows of 62 valid pixels; PW = 32: four rows of 30)
//   work item  : (N tile, patch) with T <= TMAX M tiles; their T accumulators sit side by side in TMEM (T x BN columns, two
//                sets so the epilogue of one work item overlaps the MMAs of the next)
//   K block    : 8 input channels: weights [hi, lo][tap][chunk][BN][4] arrive by ONE cp.async.bulk per K block (pre-split
//                by conv_prep_tc4_weights_kernel) and are used by all T tiles: T x 9 taps x 3 splits = 27 T MMAs per
//                36 KB (BN = 64) of weights
//   accuracy   : every accumulate of the tensor core truncates the running fp32 sum, so the error of a long reduction grows with
//                the NUMBER of MMAs chained into one accumulator (measured 1.5e-5 rel-L2 at K = 2304 with all three 3xTF32
//                terms in one accumulator).  The two small terms (lo x hi, hi x lo; 2^-11 of the result) therefore go to a
//                SEPARATE correction accumulator and only hi x hi to the main one: a third of the truncation events on
//                the large sum; the epilogue adds the two.  TMEM: columns [0, 256) main, [256, 512) correction, T x BN each.
//   hand-over  : per M tile (tfull[t] / tempty[t]): in the last K block the issuer commits after each tile's MMAs, so the
//                epilogue of tile 0 runs under the MMAs of tiles 1..T-1 and the next work item re-enters a tile's columns as
//                soon as that tile has been read out.
//   warps      : 0-3 producers (thread = patch positions; 8 coalesced channel loads, hi/lo split, 4 x 16-byte stores),
//                4-11 epilogue (tcgen05.ld, bias / activation / residual / split stores; two warps per TMEM lane quarter taking
//                alternate 32-channel blocks -- the accumulators fill all 512 TMEM columns, so the epilogue of a tile is
//                exposed until the tile is read out), 12 MMA issuer (one thread), 13 weight loader (one thread)
#pragma once
#include "tc_common.cuh"

namespace dd {

constexpr int C4_KC = 8;          // input channels per K block (= K of one tf32 MMA)
constexpr int C4_PSTAGES = 3;     // patch ring
constexpr int C4_WSTAGES = 2;     // weight ring
constexpr int C4_PROD_THREADS = 128;
constexpr int C4_EPI_WARP0 = 4, C4_EPI_WARPS = 8, C4_MMA_WARP = 12, C4_W_WARP = 13;   // two epilogue warps per TMEM lane quarter
constexpr int C4_THREADS = 448;
constexpr int C4_MAX_COUT = 1024;  // bias staged in shared memory (n_tiles * BN floats)

struct ConvTc4Args {
  ConvArgs a;
  const float* wsplit;   // [n_tile][kb][hi, lo][tap][chunk][BN][4]
  int kb_total, n_tiles;
  int bn;                // output channels per N tile = N of the MMAs (multiple of 16, <= 128)
  int wstage_bytes;      // 2 * 9 * 2 * bn * 16
  int PW, pw_shift;      // patch width (32 or 64 positions, incl. the two halo columns)
  int T;                 // M tiles per patch
  int rows_out;          // output rows per patch = T * 128 / PW
  int tiles_x, tiles_y;  // patches per image
  int total_work;        // n_tiles * B * tiles_x * tiles_y
  int np;                // patch positions (rows_out + 2) * PW
  int chunk_bytes;       // (np + 2) * 16: one lead and one tail position that only halo outputs read
  int pstage_bytes;      // 4 * chunk_bytes rounded up to 128
  int debug;             // development only (DD_TC4_DEBUG): 1 = producers skip loads / stores, 2 = no weight copies, 4 = epilogue skips its work
};

__host__ __device__ constexpr int c4_wstage_bytes(int bn) { return 2 * 9 * 2 * bn * 16; }
__host__ __device__ constexpr int c4_tmax(int bn) { return 256 / bn < 4 ? 256 / bn : 4; }   // T * bn <= 256 TMEM columns per accumulator set
constexpr int C4_TMAX = 4;
constexpr int C4_SMEM_MAX = tc::SMEM_BUDGET - (C4_MAX_COUT + 32) * 4 - 1024;   // dynamic shared memory next to the static bias array

// N tile: the output channels are split evenly into ceil(Cout / 128) tiles, each rounded up to the MMA's N granularity of 16
// (e.g. 64 -> 64, 67 -> 80, 112 -> 112, 240 -> 2 x 128, 512 -> 4 x 128)
static int c4_bn(int cout) {
  const int n_tiles = (cout + 127) / 128;
  return ((cout + n_tiles - 1) / n_tiles + 15) / 16 * 16;
}

static size_t conv_tc4_weight_bytes(int cin, int cout) {
  const int bn = c4_bn(cout);
  return (size_t)((cout + bn - 1) / bn) * ((cin + C4_KC - 1) / C4_KC) * c4_wstage_bytes(bn);
}

// OIHW fp32 weights -> per (N tile, K block) one contiguous block [hi, lo][tap][chunk][BN][4]: row n of chunk q of tap t is
// W[co = nt * BN + n][ci = kb * 8 + 4 q .. + 3][t] (zeros outside), hi = rn_tf32(w), lo = w - hi.  `transpose` builds the
// data-gradient operator (flipped taps, swapped channel roles).
__global__ void conv_prep_tc4_weights_kernel(const float* __restrict__ w, float* __restrict__ ws, int Cout_f, int Cin_f, int n_tiles,
                                             int kb_total, int BN, int transpose) {
  const int n_out = transpose ? Cin_f : Cout_f;
  const int n_in = transpose ? Cout_f : Cin_f;
  const size_t half = (size_t)9 * 2 * BN * 4;   // floats of one hi (or lo) block
  const size_t total = (size_t)n_tiles * kb_total * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 3);
    const int n = (int)((i >> 2) % BN);
    const int q = (int)((i / (4 * (size_t)BN)) & 1);
    const int tap = (int)((i / (8 * (size_t)BN)) % 9);
    const size_t blk = i / half;                 // nt * kb_total + kb
    const int kb = (int)(blk % kb_total), nt = (int)(blk / kb_total);
    const int ci = kb * C4_KC + 4 * q + j, co = nt * BN + n;
    float v = 0.f;
    if (ci < n_in && co < n_out)
      v = transpose ? __ldg(w + ((size_t)ci * Cin_f + co) * 9 + (8 - tap)) : __ldg(w + ((size_t)co * Cin_f + ci) * 9 + tap);
    const float hi = tc::tf32_rn(v);
    float* dst = ws + blk * 2 * half + (i % half);
    dst[0] = hi;
    dst[half] = v - hi;
  }
}

// K-major, no swizzle: rows 16 B apart, 8-row groups 128 B apart (SBO), second 16-byte K chunk lbo bytes further on
__device__ __forceinline__ uint64_t c4_desc(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(128u >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void st_global_f32(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int IL>
__global__ void __launch_bounds__(C4_THREADS, 1) conv_tc4_kernel(const __grid_constant__ ConvTc4Args g) {
  using namespace tc;
  constexpr int NSLOT = C4_TMAX + 1;                   // patch positions per producer thread: np <= T * 128 + 128
  constexpr int CORR_COL0 = 256;                       // TMEM columns [0, 256): hi x hi sums, [256, 512): lo x hi + hi x lo sums
  const int BN = g.bn;                                 // (T * BN <= 256)
  const int WSTAGE_BYTES = g.wstage_bytes;
  const int W_HALF = WSTAGE_BYTES / 2;                 // hi (or lo) block
  const int W_TAP = 2 * BN * 16;                       // one tap: two chunks of BN rows
  const ConvArgs& a = g.a;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const uint32_t w_base = smem_base;
  const uint32_t p_base = smem_base + C4_WSTAGES * WSTAGE_BYTES;
  const uint32_t bar_off = C4_WSTAGES * WSTAGE_BYTES + C4_PSTAGES * g.pstage_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  const uint32_t wfull = bar_base, wempty = bar_base + 16, pfull = bar_base + 32, pempty = bar_base + 56;
  const uint32_t tfull = bar_base + 80, tempty = bar_base + 112;     // one pair per M tile (TMAX <= 4)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 160);
  __shared__ __align__(16) float bias_s[C4_MAX_COUT + 32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C4_WSTAGES; ++s) mbar_init(wfull + 8 * s, 1), mbar_init(wempty + 8 * s, 1);
    for (int s = 0; s < C4_PSTAGES; ++s) mbar_init(pfull + 8 * s, C4_PROD_THREADS), mbar_init(pempty + 8 * s, 1);
    for (int s = 0; s < 4; ++s) mbar_init(tfull + 8 * s, 1), mbar_init(tempty + 8 * s, C4_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < g.n_tiles * BN + 32; i += C4_THREADS) bias_s[i] = (a.bias != nullptr && i < a.Cout) ? __ldg(a.bias + i) : 0.f;
  if (warp == C4_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < C4_PROD_THREADS) {
    // lead / tail positions of every chunk (read only by accumulator rows that are dropped): keep them finite
    for (int i = threadIdx.x; i < C4_PSTAGES * 4 * 2; i += C4_PROD_THREADS) {
      const int st = i >> 3, ck = (i >> 1) & 3, end = i & 1;
      st_shared_v4(p_base + st * g.pstage_bytes + ck * g.chunk_bytes + (end ? (g.np + 1) * 16 : 0), 0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_img = g.tiles_x * g.tiles_y;
  const int cols_out = g.PW - 2;

  if (warp < C4_EPI_WARP0) {
    // ------------------------------------------------------------------ producers: thread = patch positions ptid + 128 s
    const int ptid = threadIdx.x;
    const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
    const int C0 = a.vin.C0, Call = a.vin.C0 + a.vin.C1;
    uint32_t pit = 0;
    for (int work = blockIdx.x; work < g.total_work; work += gridDim.x) {
      const int patch = work / g.n_tiles;
      const int b = patch / tiles_img, rem = patch - b * tiles_img, ty = rem / g.tiles_x;
      const int y0 = ty * g.rows_out, x0 = (rem - ty * g.tiles_x) * cols_out;
      const float* img0 = a.vin.x0 + (size_t)b * C0 * plane0;
      const float* img1 = a.vin.x1 + (size_t)b * a.vin.C1 * plane1;   // never dereferenced when C1 == 0
      int o0[NSLOT], o1[NSLOT];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int pos = ptid + C4_PROD_THREADS * s;
        o0[s] = o1[s] = -1;
        if (pos < g.np) {
          TapEntry te;
          build_tile_map(a.vin, a.oy + y0 - 1 + (pos >> g.pw_shift), a.ox + x0 - 1 + (pos & (g.PW - 1)), te, o1[s]);
          o0[s] = te.o00;
        }
      }
      for (int kb = 0; kb < g.kb_total; ++kb, ++pit) {
        const uint32_t ps = pit % C4_PSTAGES, pph = (pit / C4_PSTAGES) & 1u;
        float v[NSLOT][C4_KC];
        if (g.debug & 1) {
          mbar_wait(pempty + 8 * ps, pph ^ 1u);
          mbar_arrive(pfull + 8 * ps);
          continue;
        }
#pragma unroll
        for (int j = 0; j < C4_KC; ++j) {
          const int ch = kb * C4_KC + j;   // CTA-uniform: which source tensor this channel comes from is a uniform select
          const bool from0 = ch < C0, live = ch < Call;
          const float* pl = from0 ? img0 + (size_t)ch * plane0 : img1 + (ptrdiff_t)(ch - C0) * (ptrdiff_t)plane1;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) {
            const int o = from0 ? o0[s] : o1[s];
            v[s][j] = (live && o >= 0) ? __ldg(pl + o) : 0.f;
          }
        }
        mbar_wait(pempty + 8 * ps, pph ^ 1u);
        const uint32_t stage = p_base + ps * g.pstage_bytes;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int pos = ptid + C4_PROD_THREADS * s;
          if (pos < g.np) {
            const uint32_t dst = stage + (uint32_t)(pos + 1) * 16u;
            float hi[C4_KC];
#pragma unroll
            for (int j = 0; j < C4_KC; ++j) hi[j] = tf32_rn(v[s][j]);
            st_shared_v4(dst, hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(dst + g.chunk_bytes, hi[4], hi[5], hi[6], hi[7]);
            st_shared_v4(dst + 2 * g.chunk_bytes, v[s][0] - hi[0], v[s][1] - hi[1], v[s][2] - hi[2], v[s][3] - hi[3]);
            st_shared_v4(dst + 3 * g.chunk_bytes, v[s][4] - hi[4], v[s][5] - hi[5], v[s][6] - hi[6], v[s][7] - hi[7]);
          }
        }
        fence_async_smem();
        mbar_arrive(pfull + 8 * ps);
      }
    }
  } else if (warp == C4_W_WARP) {
    // ------------------------------------------------------------------ weight loader: one bulk copy per K block
    if (lane == 0) {
      uint32_t wit = 0;
      for (int work = blockIdx.x; work < g.total_work; work += gridDim.x) {
        const int nt = work % g.n_tiles;
        for (int kb = 0; kb < g.kb_total; ++kb, ++wit) {
          const uint32_t ws = wit % C4_WSTAGES, wph = (wit / C4_WSTAGES) & 1u;
          mbar_wait(wempty + 8 * ws, wph ^ 1u);
          const float* src = g.wsplit + ((size_t)nt * g.kb_total + kb) * (WSTAGE_BYTES / 4);
          if (g.debug & 2) {
            mbar_arrive(wfull + 8 * ws);
            continue;
          }
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wfull + 8 * ws), "r"(WSTAGE_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_base + ws * WSTAGE_BYTES),
                       "l"(src), "r"(WSTAGE_BYTES), "r"(wfull + 8 * ws)
                       : "memory");
        }
      }
    }
    __syncwarp();
  } else if (warp == C4_MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs this code convergently (barrier waits, warp-uniform descriptor arithmetic in uniform registers);
    // only the tcgen05 instructions sit under elect.sync.  (Wrapping the loop in `if (lane == 0)` makes ptxas emit an
    // ELECT / BRA.U.ANY serialisation loop plus R2UR moves around EVERY MMA: ~100 clocks of issue per 32-clock MMA.)
    // D fp32 (bits 4-5 = 1), A / B tf32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 (17-22), M >> 4 (24-28)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // descriptor = {low word: (address >> 4) | (LBO >> 4) << 16, high word: SBO >> 4 | version 1 << 14 | layout 0 << 29}
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
    const uint32_t B_LBO16 = (uint32_t)BN << 16;   // (BN * 16 bytes) >> 4
    const uint32_t a_lbo16 = ((uint32_t)g.chunk_bytes >> 4) << 16, a_lo_delta = (2u * (uint32_t)g.chunk_bytes) >> 4;
    int tapoff[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tapoff[tap] = (tap / 3 - 1) * g.PW + (tap % 3 - 1);   // in 16-byte units
    uint32_t it = 0, tcount = 0;
    for (int work = blockIdx.x; work < g.total_work; work += gridDim.x, ++tcount) {
      const uint32_t iph = tcount & 1u;
      for (int kb = 0; kb < g.kb_total; ++kb, ++it) {
        const uint32_t ws = it % C4_WSTAGES, wph = (it / C4_WSTAGES) & 1u;
        const uint32_t ps = it % C4_PSTAGES, pph = (it / C4_PSTAGES) & 1u;
        mbar_wait(wfull + 8 * ws, wph);
        mbar_wait(pfull + 8 * ps, pph);
        tc_fence_after();
        const uint32_t b0 = (w_base + ws * WSTAGE_BYTES) >> 4;                                  // hi block, tap 0
        // patch position p lives at byte (p + 1) * 16 of its chunk; the first output position of M tile t is PW + 128 t
        const uint32_t a0 = ((p_base + ps * g.pstage_bytes) >> 4) + 1u + (uint32_t)g.PW;      // hi chunk 0, position PW
        const bool last_kb = kb == g.kb_total - 1;
        // M tiles are issued in groups of IL, round-robin inside a group, so that consecutive MMAs go to different
        // accumulators: back-to-back MMAs into ONE accumulator (N = 64: 32 clocks of work each) expose the tensor pipe's
        // read-modify-write latency (measured ~115 clocks per MMA with IL = 1).
        for (int tg = 0; tg < g.T; tg += IL) {
          if (kb == 0) {   // the previous work item's epilogue has read these tiles' columns
#pragma unroll
            for (int u = 0; u < IL; ++u)
              if (tg + u < g.T) mbar_wait(tempty + 8 * (tg + u), iph ^ 1u);
            tc_fence_after();
          }
          const uint32_t d0 = tmem_base + (uint32_t)(tg * BN);
          const uint32_t at = a0 + (uint32_t)(tg * 128);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t ah = at + (uint32_t)tapoff[tap];
              const uint64_t db_hi = desc64((b0 + tap * (W_TAP >> 4)) | B_LBO16, DESC_HI);
              const uint64_t db_lo = desc64((b0 + (W_HALF >> 4) + tap * (W_TAP >> 4)) | B_LBO16, DESC_HI);
              const uint32_t acc = (kb > 0 || tap > 0) ? 1u : 0u;
#pragma unroll
              for (int u = 0; u < IL; ++u)
                if (tg + u < g.T)
                  umma_tf32(d0 + u * BN + CORR_COL0, desc64((ah + u * 128 + a_lo_delta) | a_lbo16, DESC_HI), db_hi, idesc, acc);
#pragma unroll
              for (int u = 0; u < IL; ++u)
                if (tg + u < g.T) umma_tf32(d0 + u * BN + CORR_COL0, desc64((ah + u * 128) | a_lbo16, DESC_HI), db_lo, idesc, 1u);
#pragma unroll
              for (int u = 0; u < IL; ++u)
                if (tg + u < g.T) umma_tf32(d0 + u * BN, desc64((ah + u * 128) | a_lbo16, DESC_HI), db_hi, idesc, acc);
            }
            if (last_kb) {   // these tiles' accumulators are complete
#pragma unroll
              for (int u = 0; u < IL; ++u)
                if (tg + u < g.T) tc_commit(tfull + 8 * (tg + u));
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          tc_commit(pempty + 8 * ps);   // patch stage and weight stage may be refilled once these MMAs have read them
          tc_commit(wempty + 8 * ws);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps with the same (warp & 3) = ew share TMEM
    // lanes 32 ew .. 32 ew + 31 and take the 32-channel column blocks of their parity eh
    const int ew = warp & 3, eh = (warp - C4_EPI_WARP0) >> 2;
    const size_t HoWo = (size_t)a.Ho * a.Wo;
    uint32_t tcount = 0;
    for (int work = blockIdx.x; work < g.total_work; work += gridDim.x, ++tcount) {
      const int nt = work % g.n_tiles, patch = work / g.n_tiles;
      const int b = patch / tiles_img, rem = patch - b * tiles_img, ty = rem / g.tiles_x;
      const int y0 = ty * g.rows_out, x0 = (rem - ty * g.tiles_x) * cols_out;
      const uint32_t iph = tcount & 1u;
#pragma unroll 1
      for (int t = 0; t < g.T; ++t) {
        mbar_wait(tfull + 8 * t, iph);
        tc_fence_after();
        const int q = t * 128 + ew * 32 + lane;          // linear position relative to patch row 1, column 0
        const int r = q >> g.pw_shift, c = q & (g.PW - 1);
        const int y = y0 + r, x = x0 + c - 1;
        const bool valid = c >= 1 && c <= cols_out && y < a.Ho && x < a.Wo;
        const size_t pix = (size_t)y * a.Wo + x;
#pragma unroll 1
        for (int cb = eh; cb < ((g.debug & 4) ? 0 : (BN + 31) / 32); cb += 2) {
          const int co0 = nt * BN + cb * 32;
          if (co0 >= a.Cout) break;   // warp-uniform
          // Loads first, stores last: the output pointers are generic, so the compiler must assume that a store may alias
          // the shared-memory bias (or the residual) and would otherwise serialise load -> add -> store per channel.
          float v[32];
          {
            uint32_t rm[32], rc[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(t * BN + cb * 32);
            tmem_ld32(taddr, rm);
            tmem_ld32(taddr + CORR_COL0, rc);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + co0 + 4 * j4);
              v[4 * j4 + 0] = (__uint_as_float(rm[4 * j4 + 0]) + __uint_as_float(rc[4 * j4 + 0])) + b4.x;
              v[4 * j4 + 1] = (__uint_as_float(rm[4 * j4 + 1]) + __uint_as_float(rc[4 * j4 + 1])) + b4.y;
              v[4 * j4 + 2] = (__uint_as_float(rm[4 * j4 + 2]) + __uint_as_float(rc[4 * j4 + 2])) + b4.z;
              v[4 * j4 + 3] = (__uint_as_float(rm[4 * j4 + 3]) + __uint_as_float(rc[4 * j4 + 3])) + b4.w;
            }
          }
          // activations with the hardware exponential (ex2.approx: absolute error <= ~1.5e-7 on these ranges, far inside the
          // 1e-4 parity bound; the IEEE expf costs ~20 instructions per element on an epilogue that is not hidden)
          if (a.act == DD_ACT_ELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : __expf(v[j]) - 1.f;
          } else if (a.act == DD_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (a.act == DD_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
          }
          // destination of channel co: `out` (C_a channels) below the split, `out1` (Cout - split channels) from it on
          const int C_a = a.split > 0 ? a.split : a.Cout;
          const int n_here = min(32, min(a.Cout, (nt + 1) * BN) - co0);   // channels of this block that exist (and belong to this N tile)
          const bool in0 = co0 + n_here <= C_a, in1 = a.split > 0 && co0 >= a.split;
          if (in0 || in1) {
            // fast path (the block lies in one destination tensor): one pointer, one predicated store per channel
            float* p = in0 ? a.out + ((size_t)b * C_a + co0) * HoWo + pix
                           : (a.out1 == nullptr ? nullptr : a.out1 + ((size_t)b * (a.Cout - a.split) + (co0 - a.split)) * HoWo + pix);
            const bool st_ok = valid && p != nullptr;
            if (a.residual != nullptr) {   // (only single-tensor forward layers carry a residual; same offsets as `out`)
              const float* rp = a.residual + (p - a.out);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (st_ok && j < n_here) v[j] += __ldg(rp + (size_t)j * HoWo);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (st_ok && j < n_here) st_global_f32(p + (size_t)j * HoWo, v[j]);
          } else {
            // the split boundary falls inside this block of 32 channels (e.g. a 3-channel x0 in front of the skip tensor)
            float* base0 = a.out + (size_t)b * C_a * HoWo + pix;
            float* base1 = a.out1 == nullptr ? nullptr : a.out1 + (size_t)b * (a.Cout - a.split) * HoWo + pix;
#pragma unroll 1
            for (int j = 0; j < n_here; ++j) {
              const int co = co0 + j;   // warp-uniform
              float* d = co < C_a ? base0 + (size_t)co * HoWo : (base1 == nullptr ? nullptr : base1 + (size_t)(co - a.split) * HoWo);
              float val = v[0];
#pragma unroll
              for (int q = 1; q < 32; ++q) val = j == q ? v[q] : val;    // register array: select, do not index
              if (valid && d != nullptr) st_global_f32(d, val);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty + 8 * t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C4_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static int c4_dynamic_smem(const ConvTc4Args& g) {
  const int need = 1024 + C4_WSTAGES * g.wstage_bytes + C4_PSTAGES * g.pstage_bytes + 256;
  return need < 120 * 1024 ? 120 * 1024 : need;   // more than half of the SM: one CTA per SM, which owns all 512 TMEM columns
}

static int launch_conv_tc4(const ConvTc4Args& g, int sms, cudaStream_t st) {
  constexpr int IL = 2;   // M tiles interleaved per issue group (1, 2 and 4 measured the same)
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc4_kernel<IL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C4_SMEM_MAX));
    configured = true;
  }
  const int smem = c4_dynamic_smem(g);
  DD_REQUIRE(smem <= C4_SMEM_MAX, "conv_tc4_kernel: %d bytes of shared memory needed (bn %d, T %d)", smem, g.bn, g.T);
  conv_tc4_kernel<IL><<<g.total_work < sms ? g.total_work : sms, C4_THREADS, smem, st>>>(g);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

// Tensor cores are the default for every 3x3 layer the Winograd kernels used to take; DD_TC_CONV=0 restores those.
static bool use_tc4_conv(int ks, int cin, int cout) {
  static const char* env = getenv("DD_TC_CONV");
  static const bool off = env != nullptr && env[0] == '0', all = env != nullptr && env[0] == '2';
  if (off || ks != 3 || cout <= 16 || cin < 8) return false;
  // <= 32 output channels: an MMA with N = 32 still pays the full A-operand fetch (40 clocks for 16 clocks of math, see
  // dev/micro/mma_rate.cu) and these layers have few K blocks per work item; the Winograd CUDA-core kernel is level or
  // ahead there (32 -> 32 at 96x320: 0.51 ms vs 0.63 ms).  DD_TC_CONV=2 forces the tensor-core path for tests.
  return all || cout > 32;
}

static int run_conv_tc4(const ConvArgs& args, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose, int sms,
                        cudaStream_t st) {
  DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_tc4_kernel: bilinear up-sampling must be materialised first");
  ConvTc4Args g;
  memset(&g, 0, sizeof(g));
  g.a = args;
  const int BN = c4_bn(args.Cout);
  g.bn = BN;
  g.wstage_bytes = c4_wstage_bytes(BN);
  g.n_tiles = (args.Cout + BN - 1) / BN;
  g.kb_total = (args.Cin + C4_KC - 1) / C4_KC;
  DD_REQUIRE(g.n_tiles * BN <= C4_MAX_COUT, "conv_tc4_kernel: more than %d output channels", C4_MAX_COUT);
  const size_t wn = (size_t)g.n_tiles * g.kb_total * 9 * 2 * BN * 4;
  conv_prep_tc4_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(w_oihw, wt_buf, Cout_f, Cin_f, g.n_tiles,
                                                                                                     g.kb_total, BN, transpose ? 1 : 0);
  dd::count_launches(1);
  g.wsplit = wt_buf;
  g.PW = args.Wo + 2 <= 32 ? 32 : 64;
  g.pw_shift = g.PW == 32 ? 5 : 6;
  const int rows_per_tile = 128 / g.PW, cols_out = g.PW - 2;
  g.tiles_x = (args.Wo + cols_out - 1) / cols_out;
  // M tiles per patch: fewest (rounds over the SMs) x (clocks per work item); an MMA costs max(N / 2, 32 + N / 4) clocks
  // (A-operand fetch bound below N = 128, dev/micro/mma_rate.cu), a K block ~600 clocks of hand-over, a tile's epilogue
  // ~700 clocks per 32 channels.  Ties go to the larger T (more weight reuse).
  const int tmax = c4_tmax(BN);
  double best_cost = -1.0;
  g.T = 1;
  for (int T = 1; T <= tmax; ++T) {
    const int rows = T * rows_per_tile;
    const int np = (rows + 2) * g.PW;
    if (1024 + C4_WSTAGES * g.wstage_bytes + C4_PSTAGES * ((4 * (np + 2) * 16 + 127) / 128 * 128) + 256 > C4_SMEM_MAX) continue;
    const long items = (long)g.n_tiles * args.B * g.tiles_x * ((args.Ho + rows - 1) / rows);
    const long rounds = (items + sms - 1) / sms;
    const double mma = BN / 2 > 32 + BN / 4 ? BN / 2 : 32 + BN / 4;
    const double cost = rounds * (g.kb_total * (27.0 * T * mma + 600.0) + T * ((BN + 31) / 32) * 700.0);
    if (best_cost < 0 || cost <= best_cost) best_cost = cost, g.T = T;
  }
  g.rows_out = g.T * rows_per_tile;
  g.tiles_y = (args.Ho + g.rows_out - 1) / g.rows_out;
  g.total_work = g.n_tiles * args.B * g.tiles_x * g.tiles_y;
  g.np = (g.rows_out + 2) * g.PW;
  g.chunk_bytes = (g.np + 2) * 16;
  g.pstage_bytes = (4 * g.chunk_bytes + 127) / 128 * 128;
  static const char* dbg = getenv("DD_TC4_DEBUG");
  g.debug = dbg != nullptr ? atoi(dbg) : 0;
  return launch_conv_tc4(g, sms, st);
}

}  // namespace dd
