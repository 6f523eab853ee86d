// Development micro-benchmark: clocks per tcgen05.mma.kind::tf32 (M = 128) as a function of N, of the shared-memory
// operand layout (no swizzle with 16-byte rows as conv_tc4 uses it, or SWIZZLE_128B) and of how the A start address moves
// between instructions.  One CTA, one issuing thread, operands = zeros.   nvcc -arch=sm_100a -o bin/mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../../dynamo-depth_b200/csrc/tc_common.cuh"

namespace dd { void set_error(const char*, ...) {} void count_launches(int) {} }
using namespace dd::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// layout 0: no swizzle, rows 16 B apart, SBO 128 B, LBO = lbo bytes.  layout 2: SWIZZLE_128B, SBO 1024 B.
// move: 0 = same A every time, 1 = A start advances by 16 B x {0..8} (tap pattern), 2 = 4 different accumulators round robin
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int layout, int move, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) st_shared_v4(base + 16 * i, 0.f, 0.f, 0.f, 0.f);
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = layout == 0 ? ((128u >> 4) | (1u << 14)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t lbo16 = layout == 0 ? ((10272u >> 4) << 16) : (1u << 16);
    const uint32_t a0 = (base >> 4) + 65u, b0 = (base + 96 * 1024) >> 4;
    const uint32_t b_lbo16 = layout == 0 ? (((uint32_t)N * 16u >> 4) << 16) : (1u << 16);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int u = 0; u < 9; ++u) {
            const uint32_t a = a0 + (move == 1 ? (uint32_t)((u / 3 - 1) * 64 + (u % 3 - 1)) : 0u);
            const uint32_t d = tmem + (move == 2 ? (uint32_t)((u & 3) * N) % 512u : 0u);
            umma_tf32(d, desc64(a | lbo16, hi), desc64((b0 + (move == 1 ? u * 8u : 0u)) | b_lbo16, hi), idesc, 1u);
          }
        }
        tc_commit(bar_a);
      }
      __syncwarp();
      mbar_wait(bar_a, (uint32_t)rep & 1u);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int N>
static void run(int layout, int move, int iters, long long* d_out) {
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_rate_kernel<N><<<1, 128, 200 * 1024>>>(layout, move, iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk = 0;
  cudaMemcpy(&clk, d_out, sizeof(clk), cudaMemcpyDeviceToHost);
  printf("N=%3d layout=%s move=%d : %7.1f clk per MMA (floor %d)%s\n", N, layout == 0 ? "none " : "sw128", move, (double)clk / (9.0 * iters), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long));
  const int iters = 400;
  for (int layout = 0; layout <= 2; layout += 2)
    for (int move = 0; move <= 2; ++move) {
      run<32>(layout, move, iters, d_out);
      run<64>(layout, move, iters, d_out);
      run<128>(layout, move, iters, d_out);
      run<256>(layout, move, iters, d_out);
    }
  return 0;
}
