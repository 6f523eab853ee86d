// tcgen05 building blocks shared by the tensor-core kernels of libdynamo_b200 (linear_tc.cu: Lite-Mono linear layers,
// conv_tc.cuh: decoder convolutions as implicit GEMMs): mbarrier / fence / commit wrappers, UMMA shared-memory and
// instruction descriptors for tf32 operands of either major-ness, the fp32 -> (hi, lo) TF32 split, operand-tile
// loaders, CTA set-up (barriers + TMEM allocation) and the single-thread MMA issue loop.
//
// CTA anatomy (one persistent CTA per SM; G = 2 or 3 producer groups): warps 0..4G-1 = producer groups of 128 threads
// working on K blocks it % G (global fp32 -> registers -> hi/lo -> swizzled UMMA tiles -> fence.proxy.async -> `full`), the
// next four warps = epilogue (tcgen05.ld of the TMEM lane quarter warp % 4; `tmem_empty`), the last warp = MMA issuer (one
// thread; tcgen05.commit -> `empty` / `tmem_full`); two accumulators (2 x 256 TMEM columns) overlap the epilogue of tile i with the MMAs of
// tile i+1.  The producers are latency-bound (ncu: one instruction per ~10 clocks and warp, long-scoreboard + fixed-latency
// stalls), so what matters is K blocks in flight (groups) and instructions per chunk (TileMap).
#pragma once
#include "dd_common.cuh"

namespace dd {
namespace tc {

constexpr int BM = 128;                  // UMMA M: rows of C per tile = TMEM lanes
constexpr int BK = 32;                   // fp32 elements per K block = one 128-byte swizzle row
constexpr int GROUP_THREADS = 128;       // one producer group
// Producer groups G work on K blocks it % G.  Parity waits on `empty` are only safe while a producer is at most one phase
// behind the MMA warp; a group re-enters the ring after G K blocks, so G <= stages.  Wide tiles (BN > 128) have two stages
// of shared memory -> two groups (13 warps, 128 registers per thread); narrower ones three (17 warps, 96 registers).
__host__ __device__ constexpr int groups_for(int stages) { return stages >= 3 ? 3 : 2; }
__host__ __device__ constexpr int epi_warp0(int groups) { return 4 * groups; }   // multiple of 4: epilogue warp w reads TMEM lane quarter w % 4
// Epilogue warps ew = 4 or 8: with eight, two warps share a TMEM lane quarter and take alternate 32-column blocks -- the epilogue's
// global stores were exposed with four (profiles/r02_linear_ablation.txt); the weight-gradient contraction and narrow tiles keep four
// (the smaller register budget of a 21-warp CTA made those variants spill and run 10-30 % slower).
__host__ __device__ constexpr int mma_warp(int groups, int ew) { return 4 * groups + ew; }
__host__ __device__ constexpr int cta_threads(int groups, int ew) { return (4 * groups + ew + 1) * 32; }
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int EPI_PITCH = 36;            // floats; 16-byte aligned rows, conflict-free 128-bit accesses
__host__ __device__ constexpr int epi_bytes(int ew) { return ew * 32 * EPI_PITCH * 4; }
constexpr int MAX_STAGES = 6;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int TMEM_COLS = 512;


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol error traps (the launch fails loudly) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((spin & 1023u) == 1023u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 8000000000ll) __trap();   // ~4 s at 1.9 GHz
    }
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor (tcgen05).  Addresses / offsets in 16-byte units.
//   K-major  tile [row][32 fp32], SWIZZLE_128B (16-byte chunk ^= row % 8): 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major tile [row/32][k][32 fp32], SWIZZLE_128B_BASE32B -- the only MN-major layout of 32-bit operands (32-byte
//            chunk ^= k % 4, atom = 4 reduction steps x 128 B): groups of 32 rows mn_group_bytes apart (LBO; 4096 for a
//            32-step K block), groups of 4 reduction steps 512 B apart (SBO); one instruction (K = 8) covers two of them
template <bool MN>
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t mn_group_bytes = 4096u) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(MN ? (mn_group_bytes >> 4) : 1u) << 16;
  d |= (uint64_t)(MN ? (512u >> 4) : (1024u >> 4)) << 32;
  d |= (uint64_t)1 << 46;               // descriptor version (sm_100)
  d |= (uint64_t)(MN ? 1 : 2) << 61;    // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Round to TF32 (10 mantissa bits), nearest with ties away from zero -- what cvt.rna.tf32.f32 computes for finite
// inputs, in two integer instructions (the cvt is emulated by ~9 on sm_100a: Inf / NaN handling that operands of a
// convolution or linear layer do not need; a mantissa carry propagates into the exponent correctly).
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// hi / lo split of one 16-byte chunk into the two operand tiles (same swizzled offset in both)
__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  const float hx = tf32_rn(v.x), hy = tf32_rn(v.y), hz = tf32_rn(v.z), hw = tf32_rn(v.w);
  st_shared_v4(hi_addr, hx, hy, hz, hw);
  st_shared_v4(lo_addr, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
}

// Operand-tile geometry.  A producer thread moves NF4 16-byte chunks per tile; chunk i of thread t sits at
// (row, reduction step) = (row_t + ROW_STEP * i', k_t + K_STEP * i'') with a shared-memory offset soff_t + constant(i):
// everything that depends on the thread is derived once per kernel (init), everything that depends on i folds into
// immediates of the unrolled loops -- per chunk the K loop issues one pointer add, one load, the hi/lo split and two stores.
//   K-major  (tile ROWS x 32 floats)          : t -> row t/8 (+16 i), 16-byte column t%8          -> soff += 2048 i
//   MN-major, ROWS == 128 (32 steps x 128 rows): t -> step t/32 (+4 i), rows 4 (t%32) ..           -> soff += 512 i
//   MN-major, other ROWS                       : t -> step t/8 (+16 (i&1)), rows 32 (i>>1) + 4 (t%8) -> soff += 4096 (i>>1) + 2048 (i&1)
template <bool MN, int ROWS>
struct TileMap {
  static constexpr bool WIDE = MN && ROWS == 128;
  int row_t, k_t;
  uint32_t soff_t;
  __device__ __forceinline__ void init(int ptid) {
    if (!MN) {
      row_t = ptid >> 3;
      const int ch = ptid & 7;
      k_t = ch * 4;
      soff_t = (uint32_t)(row_t * 128 + ((ch ^ (row_t & 7)) << 4));
    } else if (WIDE) {
      const int c4 = ptid & 31;
      k_t = ptid >> 5;
      row_t = c4 * 4;
      soff_t = (uint32_t)((c4 >> 3) * 4096 + k_t * 128 + (((((c4 & 7) >> 1) ^ (k_t & 3)) << 5) | ((c4 & 1) << 4)));
    } else {
      const int c8 = ptid & 7;
      k_t = ptid >> 3;
      row_t = c8 * 4;
      soff_t = (uint32_t)(k_t * 128 + ((((c8 >> 1) ^ (k_t & 3)) << 5) | ((c8 & 1) << 4)));
    }
  }
  __device__ static __forceinline__ constexpr int row_step(int i) { return !MN ? 16 * i : (WIDE ? 0 : 32 * (i >> 1)); }
  __device__ static __forceinline__ constexpr int k_step(int i) { return !MN ? 0 : (WIDE ? 4 * i : 16 * (i & 1)); }
  __device__ static __forceinline__ constexpr uint32_t s_step(int i) {
    return !MN ? 2048u * i : (WIDE ? 512u * i : 4096u * (i >> 1) + 2048u * (i & 1));
  }
};

// rows [row0, ..) x reduction steps [k0, k0 + 32) of a row-major matrix -> registers (zeros outside rows_total x k_total)
template <bool MN, int ROWS, int NF4>
__device__ __forceinline__ void load_tile(float4 (&v)[NF4], const TileMap<MN, ROWS>& m, const float* __restrict__ src, long long ld,
                                          int row0, int rows_total, int k0, int k_total) {
  const int gr0 = row0 + m.row_t, gk0 = k0 + m.k_t;
  const float* p0 = MN ? src + (long long)gk0 * ld + gr0 : src + (long long)gr0 * ld + gk0;
#pragma unroll
  for (int i = 0; i < NF4; ++i) {
    const int rs = TileMap<MN, ROWS>::row_step(i), ks = TileMap<MN, ROWS>::k_step(i);
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr0 + rs < rows_total && gk0 + ks < k_total)
      v[i] = __ldg(reinterpret_cast<const float4*>(MN ? p0 + (long long)ks * ld + rs : p0 + (long long)rs * ld + ks));
  }
}

template <bool MN, int ROWS, int NF4>
__device__ __forceinline__ void store_tile(const float4 (&v)[NF4], const TileMap<MN, ROWS>& m, uint32_t hi_base, uint32_t lo_base) {
#pragma unroll
  for (int i = 0; i < NF4; ++i) {
    const uint32_t o = m.soff_t + TileMap<MN, ROWS>::s_step(i);
    split_store(hi_base + o, lo_base + o, v[i]);
  }
}

// ---- CTA set-up ------------------------------------------------------------------------------------
struct Cta {
  uint32_t smem_base;      // shared-window address of the 1024-byte aligned stage area
  uint8_t* smem;           // generic pointer to the same byte
  int stages;
  uint32_t full_bar, empty_bar, tfull_bar, tempty_bar;
  uint32_t tmem_base;
  float* epi_stage;        // ew x 32 x EPI_PITCH floats
};

// dynamic shared memory: [<= 1023 B alignment slack][stages x STAGE_BYTES][epi_bytes(ew)][BAR_BYTES]
__host__ __device__ constexpr int smem_bytes(int stages, int stage_bytes, int ew) { return 1024 + stages * stage_bytes + epi_bytes(ew) + BAR_BYTES; }
__host__ __device__ constexpr int stages_for(int stage_bytes, int ew) {
  return (SMEM_BUDGET - 1024 - epi_bytes(ew) - BAR_BYTES) / stage_bytes > MAX_STAGES ? MAX_STAGES
                                                                                      : (SMEM_BUDGET - 1024 - epi_bytes(ew) - BAR_BYTES) / stage_bytes;
}

__device__ __forceinline__ Cta cta_setup(uint8_t* smem_raw, int stages, int stage_bytes, int mma_warp_id, int full_count, int ew) {
  Cta c;
  const uint32_t raw_addr = smem_u32(smem_raw);
  c.smem_base = (raw_addr + 1023u) & ~1023u;   // swizzle atoms are 1024-byte aligned
  c.smem = smem_raw + (c.smem_base - raw_addr);
  c.stages = stages;
  c.epi_stage = reinterpret_cast<float*>(c.smem + (size_t)stages * stage_bytes);
  const uint32_t bar_base = c.smem_base + (uint32_t)(stages * stage_bytes) + (uint32_t)epi_bytes(ew);
  c.full_bar = bar_base, c.empty_bar = bar_base + 8 * MAX_STAGES;
  c.tfull_bar = bar_base + 16 * MAX_STAGES, c.tempty_bar = c.tfull_bar + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(c.smem + (size_t)stages * stage_bytes + epi_bytes(ew) + 16 * MAX_STAGES + 32);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(c.full_bar + 8 * s, full_count);
      mbar_init(c.empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(c.tfull_bar + 8 * b, 1);
      mbar_init(c.tempty_bar + 8 * b, ew * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == mma_warp_id) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_slot;
  return c;
}

__device__ __forceinline__ void cta_teardown(const Cta& c, int mma_warp_id) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == mma_warp_id) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- MMA issuer: called by the whole MMA warp; lane 0 issues -----------------------------------------
// Tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; tile t reduces K blocks [kb0, kb1) of split t / tiles_mn.
// one lane of the (converged) warp; always the same one, so that tcgen05.commit tracks the MMAs this lane issued
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

template <bool A_MN, bool B_MN, int BN>
__device__ __forceinline__ void mma_issue_loop(const Cta& c, int total_tiles, int tiles_mn, int kb_per_split, int kb_total) {
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  // The whole warp runs the loop convergently (barrier waits, warp-uniform descriptor arithmetic); only the tcgen05
  // instructions sit under elect.sync.  (Wrapping the loop in `if (lane == 0)` makes ptxas serialise every UTCHMMA behind an
  // ELECT / BRA.U.ANY loop with R2UR moves -- measured ~100 clocks of issue per MMA in the convolution kernel.)
  // instruction descriptor: D fp32 (bits 4-5 = 1), A / B tf32 (bits 7-9, 10-12 = 2), major-ness (15, 16), N>>3 (17-22), M>>4 (24-28)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t stages = (uint32_t)c.stages;
  uint32_t it = 0, tcount = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
    const int sp = tile / tiles_mn;
    const int kb0 = sp * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
    const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
    mbar_wait(c.tempty_bar + 8 * buf, tph ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = c.tmem_base + buf * 256u;
    for (int kb = kb0; kb < kb1; ++kb, ++it) {
      const uint32_t stage = it % stages, ph = (it / stages) & 1u;
      mbar_wait(c.full_bar + 8 * stage, ph);
      tc_fence_after();
      const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
      const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t oa = ks * (A_MN ? 1024u : 32u), ob = ks * (B_MN ? 1024u : 32u);
          const uint64_t da_hi = umma_desc<A_MN>(a_hi + oa), da_lo = umma_desc<A_MN>(a_lo + oa);
          const uint64_t db_hi = umma_desc<B_MN>(b_hi + ob), db_lo = umma_desc<B_MN>(b_lo + ob);
          umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
          umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
        }
        tc_commit(c.empty_bar + 8 * stage);   // the stage may be refilled once these MMAs have read it
        if (kb == kb1 - 1) tc_commit(c.tfull_bar + 8 * buf);       // accumulator complete
      }
      __syncwarp();
    }
    if (kb1 <= kb0) {   // (empty K range: still hand the accumulator over)
      if (elect_one()) tc_commit(c.tfull_bar + 8 * buf);
      __syncwarp();
    }
  }
}

// 32 TMEM lanes (this warp's quarter) x 32 consecutive columns -> r[0..31] of the lane's row
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc
}  // namespace dd
