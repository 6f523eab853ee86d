// Weight gradient of the 3x3 decoder convolutions on the tcgen05 tensor cores (3xTF32, fp32 accuracy):
//
//      dW[co][ci][dy][dx] = sum over images and pixels  g[co][y][x] * X[ci][y + dy][x + dx]
//
// (g = gradient at the convolution output after the activation derivative, X = the layer's VIRTUAL input: padding, nearest
// up-sampling and skip concatenation resolved by the loader; networks/layers.py:85-121 as used by depth_decoder.py:40-55,
// 99-115, motion_decoder.py:34-62, pose_decoder.py:16-37 in the reference).  The reduction index of the GEMM is the PIXEL, which
// is the contiguous direction of both NCHW operands, so both are K-major and a pixel shift along x is a shift INSIDE the 16-byte
// K chunks: the loader therefore writes the input row three times, shifted by dx = -1, 0, +1 pixels, and stacks the copies as
// rows of ONE operand; a shift by dy is a different row slot.  Per image row y of a 32-pixel column strip:
//
//   A (M = 128 rows) : [dx = -1 | 0 | +1] x 32 input channels (+ 32 don't-care rows), K = the 32 pixels of row y + dy (row slot y + dy)
//   B (N <= 80 rows) : g, output channels of the tile, K = the 32 pixels of row y
//   D[dy]            : (dx, ci) x co, one accumulator pair per dy: hi x hi in TMEM columns [0, 256), the two small 3xTF32
//                      terms in [256, 512) (the tensor core truncates every accumulate; see conv_tc4.cuh)
//
// i.e. 3 dy x 4 K steps x 3 splits = 36 MMAs per row.  Both operands use the canonical no-swizzle K-major layout
// [16-byte K chunk][row][4 pixels] (rows 16 B apart, LBO = rows * 16).  Row slots form rings (5 input rows, 4 gradient rows):
// marching down a strip, every input row is loaded and split once and used by three output rows.
//
// Work item = (input-channel tile, output-channel tile, split of the (image, strip, row-range) list); its partial result goes,
// un-transformed and fully coalesced, to a workspace slab; wgrad_tc_reduce_kernel sums the slabs in a fixed order into
// dW[co][ci][3][3] (deterministic, no atomics).  The bias gradient is left to conv_bias_grad_kernel.
//
//   warps 0-3  input-row producers, warps 4-7 gradient-row producers: ONE WARP PER ROW (rows dealt round robin), so four
//              input rows and four gradient rows are in flight per SM -- the loader is bound by global-memory latency
//   warps 8-11 epilogue, warp 12 MMA issuer (warp-uniform code, MMAs under elect.sync)
// A work item covers at most WT_MAX_ROWS strip rows: the tensor core truncates every accumulate, so the error of a partial
// sum grows with the length of its chain (measured 4.6e-5 rel-L2 with 1660 accumulates, fp32 summation of the slabs after that).
#pragma once
#include "conv_tc4.cuh"

namespace dd {

constexpr int WT_SW = 32;             // pixels of a strip row = K of a row (4 MMAs of K = 8)
constexpr int WT_CHUNKS = WT_SW / 4;  // 16-byte K chunks per row
constexpr int WT_CI = 32;             // input channels per tile (three dx copies + one zero group = M 128)
constexpr int WT_XSLOTS = 5, WT_GSLOTS = 4;   // rings: at least as many slots as producer warps (parity waits), window of 3 input rows
constexpr int WT_XROWS = 3 * WT_CI;   // 96 rows are stored; the MMA's rows 96..127 read on into the next chunk (finite data, rows dropped)
constexpr int WT_XCHUNK_BYTES = WT_XROWS * 16;
constexpr int WT_XSLOT_BYTES = 2 * WT_CHUNKS * WT_XCHUNK_BYTES;   // hi + lo: 8 chunks x 96 rows x 16 B each
constexpr int WT_THREADS = 416;
constexpr int WT_GROUP = 64;          // threads of one producer group
constexpr int WT_G_WARP0 = 4, WT_EPI_WARP0 = 8, WT_MMA_WARP = 12;
constexpr int WT_MAX_ROWS = 128;      // strip rows per work item: bounds the accumulate chain (512 per accumulator, see below)

struct WgradTcArgs {
  VirtIn vin;
  const float* g;        // (B, Cout, H, W)
  int B, H, W, Cin, Cout;
  int bn;                // output channels per tile (multiple of 16, <= 80)
  int ci_tiles, co_tiles;
  int strips;            // ceil(W / 32)
  int rows_total;        // B * strips * H  (strip rows, image-major, then strip, then y)
  int rows_per_split, splits;
  float* slabs;          // [ci_tile][co_tile][split][dy 3][bn][128]
  float* gb;             // optional (Cout,) bias gradient, zero-initialised: summed by the work items of input-channel tile 0
};

__host__ __device__ constexpr int wt_gslot_bytes(int bn) { return 2 * WT_CHUNKS * bn * 16; }

// separable source maps of the virtual input (cf. build_tile_map): row / column of the x0 and x1 planes, -1 = zero
__device__ __forceinline__ void wt_axis_map(const VirtIn& v, int p, int n_in, int& i0, int& i1) {
  i0 = i1 = -1;
  if (p < -1 || p > n_in) return;
  if (v.pad_mode == DD_PAD_REFLECT) p = reflect1(p, n_in);
  else if (p < 0 || p >= n_in) return;
  i1 = p;
  i0 = v.up0 == DD_UP_NEAREST2 ? (p >> 1) : p;
}

__global__ void __launch_bounds__(WT_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgradTcArgs a) {
  using namespace tc;
  constexpr int CORR_COL0 = 256;
  const int BN = a.bn;
  const int GSLOT_BYTES = wt_gslot_bytes(BN);
  const int G_HALF = GSLOT_BYTES / 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const uint32_t x_base = smem_base, g_base = smem_base + WT_XSLOTS * WT_XSLOT_BYTES;
  const uint32_t bar_off = WT_XSLOTS * WT_XSLOT_BYTES + WT_GSLOTS * GSLOT_BYTES;
  const uint32_t bar_base = smem_base + bar_off;
  const uint32_t xfull = bar_base, xempty = bar_base + 48, gfull = bar_base + 96, gempty = bar_base + 128, dfull = bar_base + 160;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 176);
  static_assert(WT_XSLOTS <= 6 && WT_GSLOTS <= 4, "barrier block layout");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WT_XSLOTS; ++s) mbar_init(xfull + 8 * s, 32), mbar_init(xempty + 8 * s, 1);
    for (int s = 0; s < WT_GSLOTS; ++s) mbar_init(gfull + 8 * s, 32), mbar_init(gempty + 8 * s, 1);
    mbar_init(dfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WT_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // (accumulator rows 96..127 belong to no (dx, ci): their operand rows alias the next chunk / slot, their results are never read)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item
  const int pair = blockIdx.x, split = blockIdx.y;
  const int cit = pair / a.co_tiles, cot = pair - cit * a.co_tiles;
  const int ci0 = cit * WT_CI, co0 = cot * BN;
  const int row_a = split * a.rows_per_split, row_b = min(a.rows_total, row_a + a.rows_per_split);   // strip rows [row_a, row_b)
  const int strip_rows = a.strips * a.H;   // per image

  if (warp < WT_G_WARP0) {
    // ------------------------------------------------------------------ input-row producers: ONE WARP PER ROW (warp w takes the
    // rows with sequence number = w mod 4), so that four rows' loads are in flight per SM; lane = 4-pixel group x 8 channels
    const int xq = lane & 7, cig = lane >> 3;               // channels cig, cig + 4, ..., cig + 28
    const size_t plane0 = (size_t)a.vin.H0 * a.vin.W0, plane1 = (size_t)a.vin.Hin * a.vin.Win;
    const int C0 = a.vin.C0, Call = a.vin.C0 + a.vin.C1;
    uint32_t xit = 0;               // input rows produced so far (by all four warps) = ring position
    int wb = row_a / strip_rows, wrem = row_a - wb * strip_rows, wstrip = wrem / a.H, wy = wrem - wstrip * a.H;
    int cur_img = -1, cur_strip = -1, next_x_row = 0;   // next input row (image coordinates) of the current strip
    int xo0[6], xo1[6];
    for (int sr = row_a; sr < row_b; ++sr) {
      if (wb != cur_img || wstrip != cur_strip) {   // new strip (or first row of this work item): restart the window at y - 1
        cur_img = wb, cur_strip = wstrip, next_x_row = wy - 1;
#pragma unroll
        for (int j = 0; j < 6; ++j) wt_axis_map(a.vin, wstrip * WT_SW + 4 * xq - 1 + j, a.vin.Win, xo0[j], xo1[j]);
      }
      // input rows up to y + 1 (one new row per step inside a strip, three at its start)
      for (; next_x_row <= wy + 1; ++next_x_row, ++xit) {
        if ((xit & 3u) != (uint32_t)warp) continue;
        const uint32_t slot = xit % WT_XSLOTS, ph = (xit / WT_XSLOTS) & 1u;
        int r0i, r1i;
        wt_axis_map(a.vin, next_x_row, a.vin.Hin, r0i, r1i);
        float v[8][6];
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const int ch = ci0 + cig + 4 * c8;
          const bool from0 = ch < C0, live = ch < Call && r1i >= 0;
          const float* pl = from0 ? a.vin.x0 + ((size_t)wb * C0 + ch) * plane0 + (size_t)max(r0i, 0) * a.vin.W0
                                  : a.vin.x1 + ((size_t)wb * a.vin.C1 + (ch - C0)) * plane1 + (size_t)max(r1i, 0) * a.vin.Win;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int o = from0 ? xo0[j] : xo1[j];
            v[c8][j] = (live && o >= 0) ? __ldg(pl + o) : 0.f;
          }
        }
        mbar_wait(xempty + 8 * slot, ph ^ 1u);
        const uint32_t hi_base = x_base + slot * WT_XSLOT_BYTES + xq * WT_XCHUNK_BYTES, lo_base = hi_base + WT_XSLOT_BYTES / 2;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float h[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) h[j] = tf32_rn(v[c8][j]);
#pragma unroll
          for (int d = 0; d < 3; ++d) {   // copy d holds pixels x + d - 1 .. x + d + 2: tap dx = d - 1
            const uint32_t ro = (uint32_t)((d * 32 + cig + 4 * c8) * 16);
            st_shared_v4(hi_base + ro, h[d], h[d + 1], h[d + 2], h[d + 3]);
            st_shared_v4(lo_base + ro, v[c8][d] - h[d], v[c8][d + 1] - h[d + 1], v[c8][d + 2] - h[d + 2], v[c8][d + 3] - h[d + 3]);
          }
        }
        fence_async_smem();
        mbar_arrive(xfull + 8 * slot);
      }
      if (++wy == a.H) {
        wy = 0;
        if (++wstrip == a.strips) wstrip = 0, ++wb;
      }
    }
  } else if (warp < WT_EPI_WARP0) {
    // ------------------------------------------------------------------ gradient-row producers: one warp per row as well
    const int gw = warp - WT_G_WARP0;
    const size_t gplane = (size_t)a.H * a.W;
    const bool vec = (a.W & 3) == 0;
    const bool do_bias = a.gb != nullptr && cit == 0;   // every gradient pixel passes through exactly one work item of tile 0
    float bsum[20];
#pragma unroll
    for (int q = 0; q < 20; ++q) bsum[q] = 0.f;
    for (int sr = row_a + gw; sr < row_b; sr += 4) {
      const uint32_t git = (uint32_t)(sr - row_a);
      const int b = sr / strip_rows, rem = sr - b * strip_rows, strip = rem / a.H, y = rem - strip * a.H;
      const int x0 = strip * WT_SW;
      const uint32_t slot = git % WT_GSLOTS, ph = (git / WT_GSLOTS) & 1u;
      const float* grow = a.g + ((size_t)b * a.Cout + co0) * gplane + (size_t)y * a.W + x0;
      const uint32_t hi_base = g_base + slot * GSLOT_BYTES, lo_base = hi_base + G_HALF;
      // two batches of up to 10 pieces of 16 bytes per lane (BN * 8 / 32 <= 20)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float4 gv[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) {
          const int idx = lane + 32 * (q + 10 * half), co = idx >> 3, ck = idx & 7;
          gv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (co < BN && co0 + co < a.Cout) {
            const int x = x0 + 4 * ck;
            const float* p = grow + (size_t)co * gplane + 4 * ck;
            if (vec && x + 3 < a.W) gv[q] = __ldg(reinterpret_cast<const float4*>(p));
            else {
              if (x < a.W) gv[q].x = __ldg(p);
              if (x + 1 < a.W) gv[q].y = __ldg(p + 1);
              if (x + 2 < a.W) gv[q].z = __ldg(p + 2);
              if (x + 3 < a.W) gv[q].w = __ldg(p + 3);
            }
          }
        }
        if (half == 0) mbar_wait(gempty + 8 * slot, ph ^ 1u);
#pragma unroll
        for (int q = 0; q < 10; ++q) {
          const int idx = lane + 32 * (q + 10 * half), co = idx >> 3, ck = idx & 7;
          if (co < BN) split_store(hi_base + ck * (BN * 16) + co * 16, lo_base + ck * (BN * 16) + co * 16, gv[q]);
          if (half == 0) bsum[q] += (gv[q].x + gv[q].y) + (gv[q].z + gv[q].w);
          else bsum[q + 10] += (gv[q].x + gv[q].y) + (gv[q].z + gv[q].w);
        }
      }
      fence_async_smem();
      mbar_arrive(gfull + 8 * slot);
    }
    if (do_bias) {   // lanes 8k .. 8k+7 hold the eight 4-pixel pieces of channel (lane >> 3) + 4 q: reduce, one atomic per channel
#pragma unroll
      for (int q = 0; q < 20; ++q) {
        float v = bsum[q];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        const int co = (lane >> 3) + 4 * q;
        if ((lane & 7) == 0 && co < BN && co0 + co < a.Cout) atomicAdd(a.gb + co0 + co, v);
      }
    }
  } else if (warp == WT_MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
    constexpr uint32_t A_LBO16 = ((uint32_t)WT_XCHUNK_BYTES >> 4) << 16;
    const uint32_t B_LBO16 = (uint32_t)BN << 16;
    uint32_t xit = 0, git = 0;     // input rows consumed (window start) / gradient rows consumed
    uint32_t x_lo = 0;             // ring index of input row y - 1 of the current strip row
    int cur_img = -1, cur_strip = -1;
    bool first = true;
    for (int sr = row_a; sr < row_b; ++sr, ++git) {
      const int b = sr / strip_rows, rem = sr - b * strip_rows, strip = rem / a.H;
      if (b != cur_img || strip != cur_strip) {
        // new strip: the producers restart the window, i.e. the rows loaded so far (xit of them) are followed by y-1, y, y+1
        if (cur_img >= 0) {   // release the last two rows of the previous strip's window
          if (elect_one()) {
            tc_commit(xempty + 8 * (x_lo % WT_XSLOTS));
            tc_commit(xempty + 8 * ((x_lo + 1) % WT_XSLOTS));
          }
          __syncwarp();
          xit += 2;
        }
        cur_img = b, cur_strip = strip;
        x_lo = xit;
      }
      // rows x_lo, x_lo + 1, x_lo + 2 of the ring = image rows y - 1, y, y + 1
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const uint32_t it = x_lo + d;
        mbar_wait(xfull + 8 * (it % WT_XSLOTS), (it / WT_XSLOTS) & 1u);
      }
      const uint32_t gs = git % WT_GSLOTS;
      mbar_wait(gfull + 8 * gs, (git / WT_GSLOTS) & 1u);
      tc_fence_after();
      const uint32_t b_hi = (g_base + gs * GSLOT_BYTES) >> 4, b_lo = b_hi + ((uint32_t)G_HALF >> 4);
      if (elect_one()) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const uint32_t xs = (x_lo + d) % WT_XSLOTS;
          const uint32_t a_hi = (x_base + xs * WT_XSLOT_BYTES) >> 4, a_lo = a_hi + ((WT_XSLOT_BYTES / 2) >> 4);
          const uint32_t d_main = tmem_base + (uint32_t)(d * BN), d_corr = d_main + CORR_COL0;
#pragma unroll
          for (int ks = 0; ks < WT_CHUNKS / 2; ++ks) {
            const uint32_t ao = (uint32_t)(2 * ks) * ((uint32_t)WT_XCHUNK_BYTES >> 4), bo = (uint32_t)(2 * ks) * (uint32_t)BN;   // chunk pair, in 16 B
            const uint64_t da_hi = desc64((a_hi + ao) | A_LBO16, DESC_HI), da_lo = desc64((a_lo + ao) | A_LBO16, DESC_HI);
            const uint64_t db_hi = desc64((b_hi + bo) | B_LBO16, DESC_HI), db_lo = desc64((b_lo + bo) | B_LBO16, DESC_HI);
            const uint32_t acc = (first && ks == 0) ? 0u : 1u;
            umma_tf32(d_corr, da_lo, db_hi, idesc, acc);
            umma_tf32(d_corr, da_hi, db_lo, idesc, 1u);
            umma_tf32(d_main, da_hi, db_hi, idesc, acc);
          }
        }
        tc_commit(xempty + 8 * (x_lo % WT_XSLOTS));   // row y - 1 is not needed by the next strip row
        tc_commit(gempty + 8 * gs);
      }
      __syncwarp();
      first = false;
      ++x_lo, ++xit;
    }
    if (elect_one()) tc_commit(dfull);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: partial result -> workspace slab
    const int ew = warp - WT_EPI_WARP0;
    mbar_wait(dfull, 0u);
    tc_fence_after();
    float* slab = a.slabs + (((size_t)pair * a.splits + split) * 3) * (size_t)BN * 128;
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
#pragma unroll 1
      for (int cb = 0; cb < (BN + 31) / 32; ++cb) {
        uint32_t rm[32], rc[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(d * BN + cb * 32);
        tmem_ld32(taddr, rm);
        tmem_ld32(taddr + CORR_COL0, rc);
        float* dst = slab + ((size_t)d * BN + cb * 32) * 128 + ew * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cb * 32 + j < BN) st_global_f32(dst + (size_t)j * 128, __uint_as_float(rm[j]) + __uint_as_float(rc[j]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WT_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// dW[co][ci][(dy+1)*3 + (dx+1)] = sum over splits of slab[pair(ci tile, co tile)][split][dy+1][co - co0][(dx+1)*32 + ci - ci0]
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ slabs, float* __restrict__ gw, int Cin, int Cout, int bn, int co_tiles,
                                       int splits) {
  const size_t total = (size_t)Cout * Cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9), ci = (int)((i / 9) % Cin), co = (int)(i / ((size_t)9 * Cin));
    const int d = tap / 3, dxi = tap % 3;
    const int cit = ci / WT_CI, cot = co / bn;
    const float* p = slabs + ((((size_t)(cit * co_tiles + cot) * splits) * 3 + d) * bn + (co - cot * bn)) * 128 + dxi * 32 + (ci - cit * WT_CI);
    const size_t stride = (size_t)3 * bn * 128;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += __ldg(p + (size_t)k * stride);
    gw[i] = s;
  }
}

static bool use_tc_wgrad(int ks, int cin, int cout) {
  static const char* env = getenv("DD_TC_WGRAD");
  static const bool off = env != nullptr && env[0] == '0', all = env != nullptr && env[0] == '2';
  if (off || ks != 3 || cout <= 16 || cin < 8) return false;
  // <= 32 output channels: N = 32 MMAs are bound by the A-operand fetch (dev/micro/mma_rate.cu); the Winograd weight gradient is
  // level or ahead there (32 -> 32 at 96x320: 1.46 vs 1.52 ms backward).  DD_TC_WGRAD=2 forces the tensor-core path for tests.
  return all || cout > 32;
}

static int wgrad_tc_bn(int cout) {
  const int tiles = (cout + 79) / 80;
  return ((cout + tiles - 1) / tiles + 15) / 16 * 16;
}

// shape of the launch for a layer: tiles, splits (one wave of the SMs over all work items), rows per split
static void wgrad_tc_plan(WgradTcArgs& a, int sms) {
  a.bn = wgrad_tc_bn(a.Cout);
  a.co_tiles = (a.Cout + a.bn - 1) / a.bn;
  a.ci_tiles = (a.Cin + WT_CI - 1) / WT_CI;
  a.strips = (a.W + WT_SW - 1) / WT_SW;
  a.rows_total = a.B * a.strips * a.H;
  const int pairs = a.ci_tiles * a.co_tiles;
  // whole waves of work items over the SMs, each item at most WT_MAX_ROWS strip rows long
  const long long work = (long long)a.rows_total * pairs;
  long long waves = (work + (long long)sms * WT_MAX_ROWS - 1) / ((long long)sms * WT_MAX_ROWS);
  waves = waves < 1 ? 1 : waves;
  int splits = (int)((waves * sms) / pairs);
  splits = splits < 1 ? 1 : (splits > a.rows_total ? a.rows_total : splits);
  a.rows_per_split = (a.rows_total + splits - 1) / splits;
  if (a.rows_per_split > WT_MAX_ROWS) a.rows_per_split = WT_MAX_ROWS;
  a.splits = (a.rows_total + a.rows_per_split - 1) / a.rows_per_split;
}

static size_t conv_wgrad_tc_slab_bytes(int B, int H, int W, int cin, int cout, int sms) {
  WgradTcArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B, a.H = H, a.W = W, a.Cin = cin, a.Cout = cout;
  wgrad_tc_plan(a, sms);
  return (size_t)a.ci_tiles * a.co_tiles * a.splits * 3 * a.bn * 128 * sizeof(float);
}

static int run_conv_wgrad_tc(WgradTcArgs& a, float* gw, int sms, cudaStream_t st) {
  DD_REQUIRE(a.vin.up0 != DD_UP_BILINEAR2, "conv_wgrad_tc_kernel: bilinear up-sampling must be materialised first");
  wgrad_tc_plan(a, sms);
  const int smem = 1024 + WT_XSLOTS * WT_XSLOT_BYTES + WT_GSLOTS * wt_gslot_bytes(a.bn) + 256;
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BUDGET - 1024));
    configured = true;
  }
  DD_REQUIRE(smem <= tc::SMEM_BUDGET - 1024, "conv_wgrad_tc_kernel: %d bytes of shared memory needed", smem);
  conv_wgrad_tc_kernel<<<dim3(a.ci_tiles * a.co_tiles, a.splits), WT_THREADS, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(a);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  const size_t n = (size_t)a.Cout * a.Cin * 9;
  wgrad_tc_reduce_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(a.slabs, gw, a.Cin, a.Cout, a.bn, a.co_tiles, a.splits);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // namespace dd
