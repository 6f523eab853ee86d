// Lite-Mono encoder nn.Linear contractions (reference: networks/depth_encoder.py:58-60 XCA qkv / proj,
// :197-199 and :243-245 pwconv1 / pwconv2 of the dilated-conv and LGFI blocks) on the 5th-generation tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM) at fp32 accuracy.
//
// Accuracy: every fp32 operand x is split on the fly into hi = rn_tf32(x) and lo = x - hi (exact), and each K step
// issues three MMAs  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  with fp32 accumulation in TMEM ("3xTF32"): the dropped
// A_lo*B_lo term and the hardware's truncation of lo to 10 mantissa bits are both <= 2^-21 relative per product.
// Measured against float64: 3e-7 .. 1.2e-5 of the output scale, growing with the reduction length (64 .. 1344; the
// tensor core's fp32 accumulator truncates), where torch's fp32 SIMT GEMM measures 2e-7 .. 1e-6 -- an order of
// magnitude inside the 1e-4 parity bound; tests/test_linear_gpu.py holds the kernel to 2e-5.
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
// Structure (persistent CTAs, one per SM, 13 or 17 warps; building blocks in tc_common.cuh):
//   producers   G = 3 groups (2 for tiles wider than 128 columns) of 128 threads working on K blocks it % G: global fp32 ->
//               registers -> (hi, lo) -> 128-byte-swizzled UMMA tiles in shared memory -> fence.proxy.async -> mbarrier `full`
//   MMA warp    one thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 8 per instruction, 12 per K block of 32),
//               tcgen05.commit releases the stage (`empty`) and hands the accumulator to the epilogue (`tmem_full`)
//   4 / 8 warps epilogue (8 for forward / input gradient: two per TMEM lane quarter on alternate column blocks): tcgen05.ld (32 lanes x 32 columns per warp) -> shared-memory transpose -> coalesced
//               128-byte row segments (+ bias) or atomics; two accumulators in TMEM (2 x 256 columns) so the
//               epilogue of tile i overlaps the MMAs of tile i+1
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dd {
namespace tc {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* Bpre;     // B_PRE kernels: B already split into (hi, lo) tile images [n tile][K block][hi | lo] (linear_prep_b_kernel)
  const float* bias;     // (Nc) added to every row, or NULL
  float* colsum;         // weight-gradient mode: (Mc) sums of A over the reduction (bias gradient), or NULL
  long long lda, ldb, ldc;
  int Mc, Nc, Kr;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int stages, atomic;
  int debug;             // DD_LINEAR_DEBUG ablations (timing experiments only): 1 = no global stores in the epilogue, 2 = no A loads
};

// eight epilogue warps where the epilogue is heavy and the CTA has register room: forward / input gradient with tiles wider than 128
// columns (two producer groups: 17 warps); narrow tiles (three producer groups) and the weight gradient keep four
__host__ __device__ constexpr int linear_epi_warps(bool b_pre, int nb32) { return (b_pre && nb32 >= 5) ? 8 : 4; }
template <int NB32, bool B_PRE>
__host__ __device__ constexpr int linear_groups() { return groups_for(stages_for(2 * A_TILE_BYTES + 2 * NB32 * 32 * BK * 4, linear_epi_warps(B_PRE, NB32))); }

// B operand of the forward / input-gradient contractions = the layer's weight: the same few tiles for every one of the
// M / 128 row tiles.  Splitting them in the producers cost more instructions than the activations themselves (12 of the 20
// 16-byte chunks per thread and K block at BN = 192), so they are split ONCE per call into the exact shared-memory image of
// a stage (hi tile | lo tile, swizzled as store_tile writes it) and arrive by one cp.async.bulk per K block.
template <bool B_MN, int NB32>
__global__ void __launch_bounds__(GROUP_THREADS) linear_prep_b_kernel(const float* __restrict__ B, long long ldb, int Nc, int Kr, int kb_total,
                                                                      float* __restrict__ dst) {
  constexpr int BN = NB32 * 32;
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int B_F4 = BN * BK / 4 / GROUP_THREADS;
  const int kb = blockIdx.x, nt = blockIdx.y;
  TileMap<B_MN, BN> mb;
  mb.init(threadIdx.x);
  float4 vb[B_F4];
  load_tile<B_MN, BN, B_F4>(vb, mb, B, ldb, nt * BN, Nc, kb * BK, Kr);
  char* img = reinterpret_cast<char*>(dst) + ((size_t)nt * kb_total + kb) * 2 * B_TILE_BYTES;
#pragma unroll
  for (int i = 0; i < B_F4; ++i) {
    const uint32_t o = mb.soff_t + TileMap<B_MN, BN>::s_step(i);
    const float hx = tf32_rn(vb[i].x), hy = tf32_rn(vb[i].y), hz = tf32_rn(vb[i].z), hw = tf32_rn(vb[i].w);
    *reinterpret_cast<float4*>(img + o) = make_float4(hx, hy, hz, hw);
    *reinterpret_cast<float4*>(img + B_TILE_BYTES + o) = make_float4(vb[i].x - hx, vb[i].y - hy, vb[i].z - hz, vb[i].w - hw);
  }
}

template <bool A_MN, bool B_MN, int NB32, bool B_PRE>
__global__ void __launch_bounds__(cta_threads(linear_groups<NB32, B_PRE>(), linear_epi_warps(B_PRE, NB32)), 1)
    linear_tc_kernel(const __grid_constant__ GemmArgs g) {
  constexpr int BN = NB32 * 32;
  constexpr int B_TILE_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  constexpr int G = linear_groups<NB32, B_PRE>();
  constexpr int EW = linear_epi_warps(B_PRE, NB32);
  constexpr int EPI_WARP0 = epi_warp0(G), MMA_WARP = mma_warp(G, EW);
  constexpr int A_F4 = BM * BK / 4 / GROUP_THREADS;   // 8
  constexpr int B_F4 = BN * BK / 4 / GROUP_THREADS;   // 2 * NB32

  extern __shared__ uint8_t smem_raw[];
  const Cta c = cta_setup(smem_raw, g.stages, STAGE_BYTES, MMA_WARP, GROUP_THREADS + (B_PRE ? 1 : 0), EW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = g.m_tiles * g.n_tiles;
  const int total_tiles = tiles_mn * g.splits;

  if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ producers
    const int grp = warp >> 2, ptid = threadIdx.x & (GROUP_THREADS - 1);
    const uint32_t stages = (uint32_t)c.stages, groups = G;
    TileMap<A_MN, BM> ma;
    TileMap<B_MN, BN> mb;
    ma.init(ptid), mb.init(ptid);
    if (B_PRE) {
      // (forward / input gradient: one split, kb in [0, kb_total)).  The group's K blocks are it = grp, grp + G, ...; the A
      // chunks of the NEXT block are requested before the current block is split and stored, so the global round trip of
      // block i+1 overlaps the shared-memory work of block i.
      const int kbt = g.kb_total;
      const long long total_blocks = (long long)((total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * kbt;   // of this CTA
      auto coords = [&](long long blk, int& mt, int& nt, int& kb) {
        const int tile = blockIdx.x + (int)(blk / kbt) * gridDim.x;
        kb = (int)(blk % kbt), nt = tile % g.n_tiles, mt = (tile / g.n_tiles) % g.m_tiles;
      };
      float4 va[A_F4], vn[A_F4];
      int mt, nt, kb;
      long long blk = grp;
      const int a_rows = (g.debug & 2) ? 0 : g.Mc;   // (ablation: every A chunk predicated off -> zeros)
      if (blk < total_blocks) {
        coords(blk, mt, nt, kb);
        load_tile<A_MN, BM, A_F4>(va, ma, g.A, g.lda, mt * BM, a_rows, kb * BK, g.Kr);
      }
      for (; blk < total_blocks; blk += groups) {
        const uint32_t it = (uint32_t)blk, stage = it % stages, ph = (it / stages) & 1u;
        const int nt_cur = nt, kb_cur = kb;
        if (blk + groups < total_blocks) {
          coords(blk + groups, mt, nt, kb);
          load_tile<A_MN, BM, A_F4>(vn, ma, g.A, g.lda, mt * BM, a_rows, kb * BK, g.Kr);
        }
        mbar_wait(c.empty_bar + 8 * stage, ph ^ 1u);
        const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES, b_hi = a_lo + A_TILE_BYTES;
        if (ptid == 0) {
          const char* src = reinterpret_cast<const char*>(g.Bpre) + ((size_t)nt_cur * kbt + kb_cur) * 2 * B_TILE_BYTES;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c.full_bar + 8 * stage), "r"(2 * B_TILE_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(b_hi), "l"(src),
                       "r"(2 * B_TILE_BYTES), "r"(c.full_bar + 8 * stage)
                       : "memory");
        }
        store_tile<A_MN, BM, A_F4>(va, ma, a_hi, a_lo);
        fence_async_smem();
        mbar_arrive(c.full_bar + 8 * stage);
#pragma unroll
        for (int i = 0; i < A_F4; ++i) va[i] = vn[i];
      }
    } else {
    uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % g.n_tiles, r = tile / g.n_tiles, mt = r % g.m_tiles, sp = r / g.m_tiles;
        const int kb0 = sp * g.kb_per_split, kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          if ((int)(it % groups) != grp) continue;
          const uint32_t stage = it % stages, ph = (it / stages) & 1u;
          float4 va[A_F4], vb[B_F4];
          load_tile<A_MN, BM, A_F4>(va, ma, g.A, g.lda, mt * BM, g.Mc, kb * BK, g.Kr);
          load_tile<B_MN, BN, B_F4>(vb, mb, g.B, g.ldb, nt * BN, g.Nc, kb * BK, g.Kr);
          if (A_MN) {   // bias gradient: a thread always holds the same four rows of A (TileMap<true, 128>)
  #pragma unroll
            for (int i = 0; i < A_F4; ++i) cs.x += va[i].x, cs.y += va[i].y, cs.z += va[i].z, cs.w += va[i].w;
          }
          mbar_wait(c.empty_bar + 8 * stage, ph ^ 1u);
          const uint32_t a_hi = c.smem_base + stage * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_lo + A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
          store_tile<A_MN, BM, A_F4>(va, ma, a_hi, a_lo);
          store_tile<B_MN, BN, B_F4>(vb, mb, b_hi, b_lo);
          fence_async_smem();
          mbar_arrive(c.full_bar + 8 * stage);
        }
        if (A_MN && g.colsum != nullptr && nt == 0) {
          const int row = mt * BM + ma.row_t;
          if (row < g.Mc) {   // Mc % 4 == 0
            atomicAdd(g.colsum + row, cs.x);
            atomicAdd(g.colsum + row + 1, cs.y);
            atomicAdd(g.colsum + row + 2, cs.z);
            atomicAdd(g.colsum + row + 3, cs.w);
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    mma_issue_loop<A_MN, B_MN, BN>(c, total_tiles, tiles_mn, g.kb_per_split, g.kb_total);
  } else {
    // ------------------------------------------------------------------ epilogue
    // warps with the same (warp & 3) = ew share TMEM lanes 32 ew .. 32 ew + 31 and take the 32-column blocks of their parity eh
    const int ew = warp & 3, eh = (warp - EPI_WARP0) >> 2;
    float* stg = c.epi_stage + (warp - EPI_WARP0) * 32 * EPI_PITCH;
    const uint32_t stg_addr = smem_u32(stg);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % g.n_tiles, mt = (tile / g.n_tiles) % g.m_tiles;
      const uint32_t buf = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(c.tfull_bar + 8 * buf, tph);
      tc_fence_after();
      const int row0 = mt * BM + ew * 32, col0 = nt * BN;
#pragma unroll 1
      for (int cb = eh; cb < NB32; cb += EW / 4) {
        if (col0 + cb * 32 >= g.Nc) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld32(c.tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 256u + (uint32_t)(cb * 32), r);
        // lane = row of the 32x32 block: transpose through shared memory so that a quarter-warp stores one 128-byte row segment
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stg_addr + (uint32_t)(lane * EPI_PITCH + 4 * j) * 4u, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int cc = (lane & 7) * 4, gcol = col0 + cb * 32 + cc;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias != nullptr && gcol < g.Nc) bv = __ldg(reinterpret_cast<const float4*>(g.bias + gcol));
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int row = rr * 4 + (lane >> 3), grow = row0 + row;
          const float4 v = *reinterpret_cast<const float4*>(stg + row * EPI_PITCH + cc);
          if (grow < g.Mc && gcol < g.Nc && !(g.debug & 1)) {   // Nc % 4 == 0
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
      mbar_arrive(c.tempty_bar + 8 * buf);   // accumulator drained: the MMA warp may overwrite it
    }
  }
  cta_teardown(c, MMA_WARP);
}

template <bool A_MN, bool B_MN, int NB32, bool B_PRE>
int launch_one(const GemmArgs& g, int sm_count, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * NB32 * 32 * BK * 4;
  GemmArgs a = g;
  constexpr int EW = linear_epi_warps(B_PRE, NB32);
  a.stages = stages_for(STAGE_BYTES, EW);
  // > half of the SM's shared memory in every configuration: one CTA per SM owns all 512 TMEM columns
  const int smem = smem_bytes(a.stages, STAGE_BYTES, EW);
  static bool configured = false;
  if (!configured) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<A_MN, B_MN, NB32, B_PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    configured = true;
  }
  if (B_PRE) {
    linear_prep_b_kernel<B_MN, NB32><<<dim3(a.kb_total, a.n_tiles), GROUP_THREADS, 0, st>>>(a.B, a.ldb, a.Nc, a.Kr, a.kb_total, const_cast<float*>(a.Bpre));
    dd::count_launches(1);
  }
  const int total = a.m_tiles * a.n_tiles * a.splits;
  const int grid = total < sm_count ? total : sm_count;
  linear_tc_kernel<A_MN, B_MN, NB32, B_PRE><<<grid, cta_threads(linear_groups<NB32, B_PRE>(), EW), smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(a);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

template <bool A_MN, bool B_MN, bool B_PRE>
int launch_nb(const GemmArgs& g, int nb32, int sm_count, cudaStream_t st) {
  switch (nb32) {
    case 1: return launch_one<A_MN, B_MN, 1, B_PRE>(g, sm_count, st);
    case 2: return launch_one<A_MN, B_MN, 2, B_PRE>(g, sm_count, st);
    case 3: return launch_one<A_MN, B_MN, 3, B_PRE>(g, sm_count, st);
    case 4: return launch_one<A_MN, B_MN, 4, B_PRE>(g, sm_count, st);
    case 5: return launch_one<A_MN, B_MN, 5, B_PRE>(g, sm_count, st);
    case 6: return launch_one<A_MN, B_MN, 6, B_PRE>(g, sm_count, st);
    case 7: return launch_one<A_MN, B_MN, 7, B_PRE>(g, sm_count, st);
    default: return launch_one<A_MN, B_MN, 8, B_PRE>(g, sm_count, st);
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
// N tile = 32 * nb32 columns: the narrowest of the widths 192 / 224 / 256 that wastes the fewest padded columns (a
// 192-wide tile measured 5-12 % faster than 224 / 256 at equal waste: lighter producers per K block); one tile when N <= 256
static int pick_nb32(int Nc) {
  static const int forced_nb32 = getenv("DD_LINEAR_MAX_NB32") ? atoi(getenv("DD_LINEAR_MAX_NB32")) : 0;
  const int n32 = (Nc + 31) / 32;
  int nb32 = 0, best_waste = 1 << 30;
  // (a 256-wide tile would leave a single stage next to the eight epilogue warps' staging buffers: widths up to 224)
  for (int cap = forced_nb32 > 0 ? forced_nb32 : 6; cap <= (forced_nb32 > 0 ? forced_nb32 : 7); ++cap) {
    const int tiles = (n32 + cap - 1) / cap, nb = (n32 + tiles - 1) / tiles;
    const int waste = ((n32 + nb - 1) / nb) * nb - n32;
    if (waste < best_waste) best_waste = waste, nb32 = nb;
  }
  return nb32;
}

// bytes of the pre-split B operand (weights) of a forward / input-gradient contraction with Nc output columns over Kr
size_t presplit_bytes(int Nc, int Kr) {
  const int nb32 = pick_nb32(Nc), n32 = (Nc + 31) / 32;
  const size_t n_tiles = (n32 + nb32 - 1) / nb32, kb_total = (Kr + BK - 1) / BK;
  return n_tiles * kb_total * 2 * (size_t)(nb32 * 32) * BK * 4;
}

int gemm(int mode, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, int Mc, int Nc, int Kr,
         const float* bias, float* colsum, void* workspace, size_t workspace_bytes, cudaStream_t st) {
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
  static const int debug = getenv("DD_LINEAR_DEBUG") ? atoi(getenv("DD_LINEAR_DEBUG")) : 0;
  g.debug = debug;
  const int nb32 = pick_nb32(Nc), n32 = (Nc + 31) / 32;
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
  if (mode != 2) {   // the weight operand is split once per call into the workspace
    const size_t need = presplit_bytes(Nc, Kr);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("linear: workspace too small (%zu < %zu)", workspace_bytes, need);
      return DD_ERR_WORKSPACE;
    }
    DD_REQUIRE(((uintptr_t)workspace & 15) == 0, "linear: workspace must be 16-byte aligned");
    g.Bpre = reinterpret_cast<const float*>(workspace);
  }
  if (mode == 0) return launch_nb<false, false, true>(g, nb32, sms, st);
  if (mode == 1) return launch_nb<false, true, true>(g, nb32, sms, st);
  return launch_nb<true, true, false>(g, nb32, sms, st);
}

}  // namespace tc
}  // namespace dd

extern "C" {

size_t dd_linear_workspace_bytes(int M, int K, int N) {
  (void)M;
  if (K <= 0 || N <= 0) return 0;
  const size_t f = dd::tc::presplit_bytes(N, K), d = dd::tc::presplit_bytes(K, N);   // forward: N columns over K; input gradient: K columns over N
  return f > d ? f : d;
}

int dd_linear_fwd(const float* x, const float* w, const float* bias, int M, int K, int N, float* y, void* workspace, size_t workspace_bytes,
                  void* stream) {
  DD_REQUIRE(x && w && y, "dd_linear_fwd: NULL pointer");
  return dd::tc::gemm(0, x, K, w, K, y, N, M, N, K, bias, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dd_linear_bwd(const float* x, const float* w, const float* grad_y, int M, int K, int N, float* grad_x, float* grad_w,
                  float* grad_b, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DD_REQUIRE(grad_y != nullptr, "dd_linear_bwd: grad_y is NULL");
  DD_REQUIRE(!(grad_b != nullptr && grad_w == nullptr), "dd_linear_bwd: grad_b is produced by the weight-gradient pass (grad_w must be given)");
  if (grad_x) {
    DD_REQUIRE(w != nullptr, "dd_linear_bwd: w is NULL");
    const int rc = dd::tc::gemm(1, grad_y, N, w, K, grad_x, K, M, K, N, nullptr, nullptr, workspace, workspace_bytes, st);
    if (rc != DD_OK) return rc;
  }
  if (grad_w) {
    DD_REQUIRE(x != nullptr, "dd_linear_bwd: x is NULL");
    DD_CHECK_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)N * K * sizeof(float), st));
    if (grad_b) DD_CHECK_CUDA(cudaMemsetAsync(grad_b, 0, (size_t)N * sizeof(float), st));
    const int rc = dd::tc::gemm(2, grad_y, N, x, K, grad_w, K, N, K, M, nullptr, grad_b, nullptr, 0, st);
    if (rc != DD_OK) return rc;
  }
  return DD_OK;
}

}  // extern "C"
