// Lite-Mono encoder nn.Linear contractions (reference: networks/depth_encoder.py:58-60 XCA qkv / proj,
// :197-199 and :243-245 pwconv1 / pwconv2 of the dilated-conv and LGFI blocks) on the 5th-generation tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM) at fp32 accuracy.
//
// Accuracy: every fp32 operand x is split on the fly into hi = rn_tf32(x) and lo = x - hi (exact), and each K step
// issues three MMAs  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  with fp32 accumulation in TMEM ("3xTF32"): the dropped
// A_lo*B_lo term and the hardware's truncation of lo to 10 mantissa bits are both <= 2^-21 relative per product, i.e.
// the result is as close to the exact product as torch's fp32 SIMT GEMM (tests: 1e-5 relative against float64).
//
// One kernel, three contractions C (Mc x Nc) = A_op * B_op^T over a reduction of length Kr, selected by which operand
// is "K-major" (reduction index contiguous in memory) or "MN-major" (row index contiguous):
//   forward      y  (M,N) = x  (M,K) . w (N,K)^T + bias : A = x  K-major,  B = w K-major
//   input grad   gx (M,K) = gy (M,N) . w (N,K)          : A = gy K-major,  B = w MN-major (reduction over N)
//   weight grad  gw (N,K) = gy (M,N)^T . x (M,K)        : A = gy MN-major, B = x MN-major (reduction over M, split
//                over CTAs, fp32 atomics into the zeroed gradient; the bias gradient = column sums of gy rides on
//                the A loader)
// No transposed copies are made: the UMMA shared-memory descriptors take either major-ness, and both are filled by
// the same coalesced 16-byte global loads.
//
// Structure (persistent CTAs, one per SM, 13 warps):
//   warps 0-7   producers, two groups of 128 threads working on alternate K blocks: global fp32 -> registers ->
//               (hi, lo) -> 128-byte-swizzled UMMA tiles in shared memory -> fence.proxy.async -> mbarrier `full`
//   warp  12    one thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 8 per instruction, 12 per K block of 32),
//               tcgen05.commit releases the stage (`empty`) and hands the accumulator to the epilogue (`tmem_full`)
//   warps 8-11  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> shared-memory transpose -> coalesced
//               128-byte row segments (+ bias) or atomics; two accumulators in TMEM (2 x 256 columns) so the
//               epilogue of tile i overlaps the MMAs of tile i+1
#include <string.h>

#include "dd_common.cuh"

namespace dd {
namespace tc {

constexpr int BM = 128;                  // UMMA M: rows of C per tile = TMEM lanes
constexpr int BK = 32;                   // fp32 elements per K block = one 128-byte swizzle row
constexpr int GROUP_THREADS = 128;       // one producer group
constexpr int EPI_WARP0 = 8;
constexpr int MMA_WARP = 12;
constexpr int THREADS = 13 * 32;
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int EPI_PITCH = 36;            // floats; 16-byte aligned rows, conflict-free 128-bit accesses
constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;
constexpr int MAX_STAGES = 6;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int TMEM_COLS = 512;

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;     // (Nc) added to every row, or NULL
  float* colsum;         // weight-gradient mode: (Mc) sums of A over the reduction (bias gradient), or NULL
  long long lda, ldb, ldc;
  int Mc, Nc, Kr;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int stages, atomic;
};

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
//            chunk ^= k % 4, atom = 4 reduction steps x 128 B): groups of 32 rows 4096 B apart (LBO), groups of 4
//            reduction steps 512 B apart (SBO); one instruction (K = 8) covers two of them
template <bool MN>
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(MN ? (4096u >> 4) : 1u) << 16;
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

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// hi / lo split of one 16-byte chunk into the two operand tiles (same swizzled offset in both)
__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  const float hx = tf32_rn(v.x), hy = tf32_rn(v.y), hz = tf32_rn(v.z), hw = tf32_rn(v.w);
  st_shared_v4(hi_addr, hx, hy, hz, hw);
  st_shared_v4(lo_addr, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
}

// Per-thread chunk c (of NF4) of an operand tile: where it comes from and where it goes.
//   K-major : tile = ROWS x 32 floats, chunk f -> row f/8, 16-byte column f%8
//   MN-major: tile = 32 reduction steps x ROWS, chunk f -> step f/(ROWS/4), rows 4*(f % (ROWS/4)) ..+3
template <bool MN, int ROWS>
struct TileMap {
  static constexpr int F4_PER_K = ROWS / 4;
  __device__ static __forceinline__ void decode(int f, int& row, int& kk, uint32_t& soff) {
    if (MN) {
      kk = f / F4_PER_K;
      const int c4 = f - kk * F4_PER_K;
      row = c4 * 4;
      soff = (uint32_t)((c4 >> 3) * 4096 + kk * 128 + (((((c4 & 7) >> 1) ^ (kk & 3)) << 5) | ((c4 & 1) << 4)));
    } else {
      row = f >> 3;
      const int ch = f & 7;
      kk = ch * 4;
      soff = (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4));
    }
  }
};

template <bool MN, int ROWS, int NF4>
__device__ __forceinline__ void load_tile(float4 (&v)[NF4], const float* __restrict__ src, long long ld, int row0, int rows_total,
                                          int k0, int k_total, int ptid) {
#pragma unroll
  for (int i = 0; i < NF4; ++i) {
    int row, kk;
    uint32_t soff;
    TileMap<MN, ROWS>::decode(ptid + GROUP_THREADS * i, row, kk, soff);
    const int gr = row0 + row, gk = k0 + kk;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < rows_total && gk < k_total) {
      const float* p = MN ? src + (long long)gk * ld + gr : src + (long long)gr * ld + gk;
      v[i] = __ldg(reinterpret_cast<const float4*>(p));
    }
  }
}

template <bool MN, int ROWS, int NF4>
__device__ __forceinline__ void store_tile(const float4 (&v)[NF4], uint32_t hi_base, uint32_t lo_base, int ptid) {
#pragma unroll
  for (int i = 0; i < NF4; ++i) {
    int row, kk;
    uint32_t soff;
    TileMap<MN, ROWS>::decode(ptid + GROUP_THREADS * i, row, kk, soff);
    split_store(hi_base + soff, lo_base + soff, v[i]);
  }
}

template <bool A_MN, bool B_MN, int NB32>
__global__ void __launch_bounds__(THREADS, 1) linear_tc_kernel(const __grid_constant__ GemmArgs g) {
  constexpr int BN = NB32 * 32;
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  constexpr int A_F4 = BM * BK / 4 / GROUP_THREADS;   // 8
  constexpr int B_F4 = BN * BK / 4 / GROUP_THREADS;   // 2 * NB32

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;   // swizzle atoms are 1024-byte aligned
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const int stages = g.stages;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)stages * STAGE_BYTES);
  const uint32_t bar_base = smem_base + (uint32_t)stages * STAGE_BYTES + EPI_BYTES;
  const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * MAX_STAGES;
  const uint32_t tfull_bar = bar_base + 16 * MAX_STAGES, tempty_bar = tfull_bar + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (size_t)stages * STAGE_BYTES + EPI_BYTES + 16 * MAX_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar + 8 * s, GROUP_THREADS);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar + 8 * b, 1);
      mbar_init(tempty_bar + 8 * b, 4 * 32);
    }
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

  const int tiles_mn = g.m_tiles * g.n_tiles;
  const int total_tiles = tiles_mn * g.splits;

  if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ producers
    const int grp = warp >> 2, ptid = threadIdx.x & (GROUP_THREADS - 1);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % g.n_tiles, r = tile / g.n_tiles, mt = r % g.m_tiles, sp = r / g.m_tiles;
      const int kb0 = sp * g.kb_per_split, kb1 = min(g.kb_total, kb0 + g.kb_per_split);
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        if ((int)(it & 1u) != grp) continue;
        const uint32_t stage = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
        float4 va[A_F4], vb[B_F4];
        load_tile<A_MN, BM, A_F4>(va, g.A, g.lda, mt * BM, g.Mc, kb * BK, g.Kr, ptid);
        load_tile<B_MN, BN, B_F4>(vb, g.B, g.ldb, nt * BN, g.Nc, kb * BK, g.Kr, ptid);
        if (A_MN) {   // bias gradient: a thread always holds the same four rows of A (GROUP_THREADS % (BM/4) == 0)
#pragma unroll
          for (int i = 0; i < A_F4; ++i) cs.x += va[i].x, cs.y += va[i].y, cs.z += va[i].z, cs.w += va[i].w;
        }
        mbar_wait(empty_bar + 8 * stage, ph ^ 1u);
        const uint32_t a_hi = smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
        const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
        store_tile<A_MN, BM, A_F4>(va, a_hi, a_lo, ptid);
        store_tile<B_MN, BN, B_F4>(vb, b_hi, b_lo, ptid);
        fence_async_smem();
        mbar_arrive(full_bar + 8 * stage);
      }
      if (A_MN && g.colsum != nullptr && nt == 0) {
        const int row = mt * BM + (ptid & (BM / 4 - 1)) * 4;
        if (row < g.Mc) {   // Mc % 4 == 0
          atomicAdd(g.colsum + row, cs.x);
          atomicAdd(g.colsum + row + 1, cs.y);
          atomicAdd(g.colsum + row + 2, cs.z);
          atomicAdd(g.colsum + row + 3, cs.w);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      // instruction descriptor: D fp32 (bits 4-5 = 1), A / B tf32 (bits 7-9, 10-12 = 2), major-ness (15, 16), N>>3 (17-22), M>>4 (24-28)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int sp = tile / tiles_mn;
        const int kb0 = sp * g.kb_per_split, kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar + 8 * buf, tph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t stage = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
          mbar_wait(full_bar + 8 * stage, ph);
          tc_fence_after();
          const uint32_t a_hi = smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t oa = ks * (A_MN ? 1024u : 32u), ob = ks * (B_MN ? 1024u : 32u);
            const uint64_t da_hi = umma_desc<A_MN>(a_hi + oa), da_lo = umma_desc<A_MN>(a_lo + oa);
            const uint64_t db_hi = umma_desc<B_MN>(b_hi + ob), db_lo = umma_desc<B_MN>(b_lo + ob);
            umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
            umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
            umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
          }
          tc_commit(empty_bar + 8 * stage);   // the stage may be refilled once these MMAs have read it
        }
        tc_commit(tfull_bar + 8 * buf);       // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - EPI_WARP0;   // == warp % 4: the TMEM lane quarter this warp may read
    float* stg = epi_stage + ew * 32 * EPI_PITCH;
    const uint32_t stg_addr = smem_u32(stg);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % g.n_tiles, mt = (tile / g.n_tiles) % g.m_tiles;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(tfull_bar + 8 * buf, tph);
      tc_fence_after();
      const int row0 = mt * BM + ew * 32, col0 = nt * BN;
#pragma unroll 1
      for (int c = 0; c < NB32; ++c) {
        if (col0 + c * 32 >= g.Nc) break;   // warp-uniform
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 256u + (uint32_t)(c * 32);
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
        // lane = row of the 32x32 block: transpose through shared memory so that a quarter-warp stores one 128-byte row segment
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stg_addr + (uint32_t)(lane * EPI_PITCH + 4 * j) * 4u, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int cc = (lane & 7) * 4, gcol = col0 + c * 32 + cc;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias != nullptr && gcol < g.Nc) bv = __ldg(reinterpret_cast<const float4*>(g.bias + gcol));
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int row = rr * 4 + (lane >> 3), grow = row0 + row;
          const float4 v = *reinterpret_cast<const float4*>(stg + row * EPI_PITCH + cc);
          if (grow < g.Mc && gcol < g.Nc) {   // Nc % 4 == 0
            float* dst = g.C + (long long)grow * g.ldc + gcol;
            if (g.atomic) {
              atomicAdd(dst, v.x), atomicAdd(dst + 1, v.y), atomicAdd(dst + 2, v.z), atomicAdd(dst + 3, v.w);
            } else {
              *reinterpret_cast<float4*>(dst) = make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * buf);   // accumulator drained: the MMA warp may overwrite it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <bool A_MN, bool B_MN, int NB32>
int launch_one(const GemmArgs& g, int sm_count, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * NB32 * 32 * BK * 4;
  GemmArgs a = g;
  int stages = (SMEM_BUDGET - 1024 - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;
  a.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
  // > half of the SM's shared memory in every configuration: one CTA per SM owns all 512 TMEM columns
  const int smem = 1024 + a.stages * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<A_MN, B_MN, NB32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    configured = true;
  }
  const int total = a.m_tiles * a.n_tiles * a.splits;
  const int grid = total < sm_count ? total : sm_count;
  linear_tc_kernel<A_MN, B_MN, NB32><<<grid, THREADS, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(a);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

template <bool A_MN, bool B_MN>
int launch_nb(const GemmArgs& g, int nb32, int sm_count, cudaStream_t st) {
  switch (nb32) {
    case 1: return launch_one<A_MN, B_MN, 1>(g, sm_count, st);
    case 2: return launch_one<A_MN, B_MN, 2>(g, sm_count, st);
    case 3: return launch_one<A_MN, B_MN, 3>(g, sm_count, st);
    case 4: return launch_one<A_MN, B_MN, 4>(g, sm_count, st);
    case 5: return launch_one<A_MN, B_MN, 5>(g, sm_count, st);
    case 6: return launch_one<A_MN, B_MN, 6>(g, sm_count, st);
    case 7: return launch_one<A_MN, B_MN, 7>(g, sm_count, st);
    default: return launch_one<A_MN, B_MN, 8>(g, sm_count, st);
  }
}

static int device_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

// mode 0: forward, 1: input gradient, 2: weight gradient (split reduction + atomics)
int gemm(int mode, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, int Mc, int Nc, int Kr,
         const float* bias, float* colsum, cudaStream_t st) {
  DD_REQUIRE(Mc > 0 && Nc > 0 && Kr > 0, "linear: empty problem %d x %d x %d", Mc, Nc, Kr);
  // 16-byte accesses: along the reduction for K-major operands, along the rows for MN-major operands and for C
  DD_REQUIRE(Nc % 4 == 0 && (mode == 2 ? Mc % 4 == 0 : Kr % 4 == 0), "linear: feature dimensions must be multiples of 4 (got %d x %d over %d)", Mc, Nc, Kr);
  DD_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "linear: leading dimensions must be multiples of 4");
  DD_REQUIRE(((uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)bias) % 16 == 0, "linear: pointers must be 16-byte aligned");
  const int sms = device_sms();
  DD_REQUIRE(sms > 0, "linear: no CUDA device");
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A, g.B = B, g.C = C, g.bias = bias, g.colsum = colsum;
  g.lda = lda, g.ldb = ldb, g.ldc = ldc, g.Mc = Mc, g.Nc = Nc, g.Kr = Kr;
  const int n32 = (Nc + 31) / 32;
  g.n_tiles = (n32 + 7) / 8;
  const int nb32 = (n32 + g.n_tiles - 1) / g.n_tiles;
  g.n_tiles = (n32 + nb32 - 1) / nb32;
  g.m_tiles = (Mc + BM - 1) / BM;
  g.kb_total = (Kr + BK - 1) / BK;
  g.splits = 1, g.kb_per_split = g.kb_total, g.atomic = 0;
  if (mode == 2) {
    const int tiles_mn = g.m_tiles * g.n_tiles;
    int splits = sms / tiles_mn;
    if (splits > g.kb_total / 4) splits = g.kb_total / 4;
    if (splits < 1) splits = 1;
    g.kb_per_split = (g.kb_total + splits - 1) / splits;
    g.splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
    g.atomic = 1;
  }
  if (mode == 0) return launch_nb<false, false>(g, nb32, sms, st);
  if (mode == 1) return launch_nb<false, true>(g, nb32, sms, st);
  return launch_nb<true, true>(g, nb32, sms, st);
}

}  // namespace tc
}  // namespace dd

extern "C" {

int dd_linear_fwd(const float* x, const float* w, const float* bias, int M, int K, int N, float* y, void* stream) {
  DD_REQUIRE(x && w && y, "dd_linear_fwd: NULL pointer");
  return dd::tc::gemm(0, x, K, w, K, y, N, M, N, K, bias, nullptr, (cudaStream_t)stream);
}

int dd_linear_bwd(const float* x, const float* w, const float* grad_y, int M, int K, int N, float* grad_x, float* grad_w,
                  float* grad_b, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DD_REQUIRE(grad_y != nullptr, "dd_linear_bwd: grad_y is NULL");
  DD_REQUIRE(!(grad_b != nullptr && grad_w == nullptr), "dd_linear_bwd: grad_b is produced by the weight-gradient pass (grad_w must be given)");
  if (grad_x) {
    DD_REQUIRE(w != nullptr, "dd_linear_bwd: w is NULL");
    const int rc = dd::tc::gemm(1, grad_y, N, w, K, grad_x, K, M, K, N, nullptr, nullptr, st);
    if (rc != DD_OK) return rc;
  }
  if (grad_w) {
    DD_REQUIRE(x != nullptr, "dd_linear_bwd: x is NULL");
    DD_CHECK_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)N * K * sizeof(float), st));
    if (grad_b) DD_CHECK_CUDA(cudaMemsetAsync(grad_b, 0, (size_t)N * sizeof(float), st));
    const int rc = dd::tc::gemm(2, grad_y, N, x, K, grad_w, K, N, K, M, nullptr, grad_b, st);
    if (rc != DD_OK) return rc;
  }
  return DD_OK;
}

}  // extern "C"
