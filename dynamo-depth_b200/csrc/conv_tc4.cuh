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
//   warps      : 0-3 producers (thread = patch positions; 8 coalesced channel loads, hi/lo split, 4 x 16-byte stores),
//                4-7 epilogue (tcgen05.ld, bias / activation / residual / split stores through emit_output),
//                8 MMA issuer (one thread), 9 weight loader (one thread)
#pragma once
#include "tc_common.cuh"

namespace dd {

constexpr int C4_KC = 8;          // input channels per K block (= K of one tf32 MMA)
constexpr int C4_PSTAGES = 3;     // patch ring
constexpr int C4_WSTAGES = 2;     // weight ring
constexpr int C4_PROD_THREADS = 128;
constexpr int C4_EPI_WARP0 = 4, C4_MMA_WARP = 8, C4_W_WARP = 9;
constexpr int C4_THREADS = 320;

struct ConvTc4Args {
  ConvArgs a;
  const float* wsplit;   // [n_tile][kb][hi, lo][tap][chunk][BN][4]
  int kb_total, n_tiles;
  int PW, pw_shift;      // patch width (32 or 64 positions, incl. the two halo columns)
  int T;                 // M tiles per patch
  int rows_out;          // output rows per patch = T * 128 / PW
  int tiles_x, tiles_y;  // patches per image
  int total_work;        // n_tiles * B * tiles_x * tiles_y
  int np;                // patch positions (rows_out + 2) * PW
  int chunk_bytes;       // (np + 2) * 16: one lead and one tail position that only halo outputs read
  int pstage_bytes;      // 4 * chunk_bytes rounded up to 128
};

__host__ __device__ constexpr int c4_wstage_bytes(int bn) { return 2 * 9 * 2 * bn * 16; }
__host__ __device__ constexpr int c4_tmax(int bn) { return bn <= 64 ? 4 : 2; }
__host__ __device__ constexpr int c4_pstage_max_bytes(int bn) { return (4 * (c4_tmax(bn) * 128 + 2 * 64 + 2) * 16 + 127) / 128 * 128; }
__host__ __device__ constexpr int c4_smem_bytes(int bn) {
  return 1024 + C4_WSTAGES * c4_wstage_bytes(bn) + C4_PSTAGES * c4_pstage_max_bytes(bn) + 256;
}

static size_t conv_tc4_weight_bytes(int cin, int cout) {
  const int bn = cout <= 32 ? 32 : (cout <= 64 ? 64 : 128);
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

template <int BN>
__global__ void __launch_bounds__(C4_THREADS, 1) conv_tc4_kernel(const __grid_constant__ ConvTc4Args g) {
  using namespace tc;
  constexpr int TMAX = c4_tmax(BN);
  constexpr int NSLOT = TMAX + 1;                      // patch positions per producer thread: np <= TMAX * 128 + 128
  constexpr int WSTAGE_BYTES = c4_wstage_bytes(BN);
  constexpr int W_HALF = WSTAGE_BYTES / 2;             // hi (or lo) block
  constexpr int W_TAP = 2 * BN * 16;                   // one tap: two chunks of BN rows
  constexpr int SET_COLS = 256;                        // TMEM columns of one accumulator set (TMAX * BN <= 256)
  static_assert(TMAX * BN <= SET_COLS, "accumulator set");
  const ConvArgs& a = g.a;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const uint32_t w_base = smem_base;
  const uint32_t p_base = smem_base + C4_WSTAGES * WSTAGE_BYTES;
  const uint32_t bar_off = C4_WSTAGES * WSTAGE_BYTES + C4_PSTAGES * c4_pstage_max_bytes(BN);
  const uint32_t bar_base = smem_base + bar_off;
  const uint32_t wfull = bar_base, wempty = bar_base + 16, pfull = bar_base + 32, pempty = bar_base + 56;
  const uint32_t tfull = bar_base + 80, tempty = bar_base + 96;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C4_WSTAGES; ++s) mbar_init(wfull + 8 * s, 1), mbar_init(wempty + 8 * s, 1);
    for (int s = 0; s < C4_PSTAGES; ++s) mbar_init(pfull + 8 * s, C4_PROD_THREADS), mbar_init(pempty + 8 * s, 1);
    for (int s = 0; s < 2; ++s) mbar_init(tfull + 8 * s, 1), mbar_init(tempty + 8 * s, 4 * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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
#pragma unroll
        for (int j = 0; j < C4_KC; ++j) {
          const int ch = kb * C4_KC + j;   // CTA-uniform
          const float* p0 = img0 + (size_t)ch * plane0;
          const float* p1 = img1 + (ptrdiff_t)(ch - C0) * (ptrdiff_t)plane1;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) {
            float x = 0.f;
            if (ch < C0) {
              if (o0[s] >= 0) x = __ldg(p0 + o0[s]);
            } else if (ch < Call) {
              if (o1[s] >= 0) x = __ldg(p1 + o1[s]);
            }
            v[s][j] = x;
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
    if (lane == 0) {
      // D fp32 (bits 4-5 = 1), A / B tf32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 (17-22), M >> 4 (24-28)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)g.chunk_bytes, a_lo_off = 2u * (uint32_t)g.chunk_bytes;
      uint32_t it = 0, tcount = 0;
      for (int work = blockIdx.x; work < g.total_work; work += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
        mbar_wait(tempty + 8 * buf, tph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < g.kb_total; ++kb, ++it) {
          const uint32_t ws = it % C4_WSTAGES, wph = (it / C4_WSTAGES) & 1u;
          const uint32_t ps = it % C4_PSTAGES, pph = (it / C4_PSTAGES) & 1u;
          mbar_wait(wfull + 8 * ws, wph);
          mbar_wait(pfull + 8 * ps, pph);
          tc_fence_after();
          const uint32_t b_hi = w_base + ws * WSTAGE_BYTES, b_lo = b_hi + W_HALF;
          // patch position p lives at byte (p + 1) * 16 of its chunk; the first output position of M tile t is PW + 128 t
          const uint32_t a_hi0 = p_base + ps * g.pstage_bytes + (uint32_t)(1 + g.PW) * 16u;
          for (int t = 0; t < g.T; ++t) {
            const uint32_t d_tmem = tmem_base + buf * SET_COLS + (uint32_t)(t * BN);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              const uint32_t a_hi = a_hi0 + (uint32_t)((t * 128 + dy * g.PW + dx) * 16);
              const uint64_t da_hi = c4_desc(a_hi, a_lbo), da_lo = c4_desc(a_hi + a_lo_off, a_lbo);
              const uint64_t db_hi = c4_desc(b_hi + tap * W_TAP, BN * 16), db_lo = c4_desc(b_lo + tap * W_TAP, BN * 16);
              umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > 0 || tap > 0) ? 1u : 0u);
              umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
            }
          }
          tc_commit(pempty + 8 * ps);   // patch stage and weight stage may be refilled once these MMAs have read them
          tc_commit(wempty + 8 * ws);
        }
        tc_commit(tfull + 8 * buf);     // all T accumulators of this work item complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: warp ew owns TMEM lanes 32 ew .. 32 ew + 31
    const int ew = warp - C4_EPI_WARP0;
    uint32_t tcount = 0;
    for (int work = blockIdx.x; work < g.total_work; work += gridDim.x, ++tcount) {
      const int nt = work % g.n_tiles, patch = work / g.n_tiles;
      const int b = patch / tiles_img, rem = patch - b * tiles_img, ty = rem / g.tiles_x;
      const int y0 = ty * g.rows_out, x0 = (rem - ty * g.tiles_x) * cols_out;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(tfull + 8 * buf, tph);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < g.T; ++t) {
        const int q = t * 128 + ew * 32 + lane;          // linear position relative to patch row 1, column 0
        const int r = q >> g.pw_shift, c = q & (g.PW - 1);
        const bool keep = c >= 1 && c <= cols_out;
        const int y = y0 + r, x = keep ? x0 + c - 1 : (1 << 28);   // emit_output drops x >= Wo
#pragma unroll 1
        for (int cb = 0; cb < BN / 32; ++cb) {
          const int co0 = nt * BN + cb * 32;
          if (co0 >= a.Cout) break;   // warp-uniform
          uint32_t rr[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * SET_COLS + (uint32_t)(t * BN + cb * 32), rr);
#pragma unroll
          for (int j = 0; j < 32; ++j) emit_output(a, b, co0 + j, y, x, __uint_as_float(rr[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C4_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN>
static int launch_conv_tc4(const ConvTc4Args& g, int sms, cudaStream_t st) {
  constexpr int SMEM = c4_smem_bytes(BN);
  static_assert(SMEM <= tc::SMEM_BUDGET, "conv_tc4: shared memory budget");
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc4_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BUDGET));
    configured = true;
  }
  // more than half of the SM's shared memory in every configuration: one CTA per SM, which owns all 512 TMEM columns
  conv_tc4_kernel<BN><<<g.total_work < sms ? g.total_work : sms, C4_THREADS, SMEM < 120 * 1024 ? 120 * 1024 : SMEM, st>>>(g);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

// Tensor cores are the default for every 3x3 layer the Winograd kernels used to take; DD_TC_CONV=0 restores those.
static bool use_tc4_conv(int ks, int cin, int cout) {
  static const char* env = getenv("DD_TC_CONV");
  static const bool off = env != nullptr && env[0] == '0';
  return !off && ks == 3 && cout > 16 && cin >= 8;
}

static int run_conv_tc4(const ConvArgs& args, float* wt_buf, const float* w_oihw, int Cout_f, int Cin_f, bool transpose, int sms,
                        cudaStream_t st) {
  DD_REQUIRE(args.vin.up0 != DD_UP_BILINEAR2, "conv_tc4_kernel: bilinear up-sampling must be materialised first");
  ConvTc4Args g;
  memset(&g, 0, sizeof(g));
  g.a = args;
  const int BN = args.Cout <= 32 ? 32 : (args.Cout <= 64 ? 64 : 128);
  g.n_tiles = (args.Cout + BN - 1) / BN;
  g.kb_total = (args.Cin + C4_KC - 1) / C4_KC;
  const size_t wn = (size_t)g.n_tiles * g.kb_total * 9 * 2 * BN * 4;
  conv_prep_tc4_weights_kernel<<<(int)((wn + 255) / 256 < 592 ? (wn + 255) / 256 : 592), 256, 0, st>>>(w_oihw, wt_buf, Cout_f, Cin_f, g.n_tiles,
                                                                                                     g.kb_total, BN, transpose ? 1 : 0);
  dd::count_launches(1);
  g.wsplit = wt_buf;
  g.PW = args.Wo + 2 <= 32 ? 32 : 64;
  g.pw_shift = g.PW == 32 ? 5 : 6;
  const int rows_per_tile = 128 / g.PW, cols_out = g.PW - 2;
  g.tiles_x = (args.Wo + cols_out - 1) / cols_out;
  // M tiles per patch: the one with the fewest (rounds over the SMs) x (work per item), ties to the larger (more weight reuse)
  const int tmax = c4_tmax(BN);
  long best_cost = -1;
  for (int T = 1; T <= tmax; ++T) {
    const int rows = T * rows_per_tile;
    const long items = (long)g.n_tiles * args.B * g.tiles_x * ((args.Ho + rows - 1) / rows);
    const long rounds = (items + sms - 1) / sms;
    const long cost = rounds * (27L * T * BN / 2 + 256);   // MMA clocks per K block + a fixed per-K-block overhead
    if (best_cost < 0 || cost <= best_cost) best_cost = cost, g.T = T;
  }
  g.rows_out = g.T * rows_per_tile;
  g.tiles_y = (args.Ho + g.rows_out - 1) / g.rows_out;
  g.total_work = g.n_tiles * args.B * g.tiles_x * g.tiles_y;
  g.np = (g.rows_out + 2) * g.PW;
  g.chunk_bytes = (g.np + 2) * 16;
  g.pstage_bytes = (4 * g.chunk_bytes + 127) / 128 * 128;
  if (BN == 32) return launch_conv_tc4<32>(g, sms, st);
  if (BN == 64) return launch_conv_tc4<64>(g, sms, st);
  return launch_conv_tc4<128>(g, sms, st);
}

}  // namespace dd
