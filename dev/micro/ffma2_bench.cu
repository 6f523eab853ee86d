// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ffma2_bench dev/micro/ffma2_bench.cu && ./gpurun_out/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%0,%1};\n"
               " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0,%1}, rc;}"
               : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float a = a0 + threadIdx.x * 1e-6f, b = b0, a2 = a0 * 0.5f, b2 = b0 * 1.01f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) ffma2(acc[i], acc[i + 1], a, a2, b, b2);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          acc[i] = fmaf(a, b, acc[i]);
          acc[i + 1] = fmaf(a2, b2, acc[i + 1]);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  const int blocks = sms * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t s, e;
  cudaEventCreate(&s), cudaEventCreate(&e);
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(s);
      if (packed) k<1><<<blocks, 256>>>(out, 1.0001f, 0.9999f, iters);
      else k<0><<<blocks, 256>>>(out, 1.0001f, 0.9999f, iters);
      cudaEventRecord(e);
      cudaEventSynchronize(e);
      float ms;
      cudaEventElapsedTime(&ms, s, e);
      const double fl = 2.0 * blocks * 256.0 * iters * 4 * 16;
      if (rep == 2) printf("%s: %.3f ms  %.1f TFLOP/s\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar ", ms, fl / ms / 1e9);
    }
  }
  return 0;
}
