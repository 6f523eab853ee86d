// Layout glue of the Lite-Mono encoder blocks (reference: networks/depth_encoder.py:204-221 DilatedConv.forward,
// :252-279 LGFI.forward).  The reference permutes (N,C,H,W) <-> (N,H,W,C) around the point-wise MLP and then computes
//   x = input + drop_path(gamma * x).permute(0, 3, 1, 2)
// which PyTorch runs as strided element-wise kernels (copy, multiply, add: ~20 ms of the bs32 training step).  Here each
// of those is one shared-memory tiled transpose that reads and writes 128-byte segments:
//   dd_nchw_to_nhwc / dd_nhwc_to_nchw : the permute(0,2,3,1).contiguous() in front of pwconv1 and its gradient
//   dd_block_tail_fwd : out (N,C,HW) = x + scale_b * gamma_c * y (N,HW,C)     [layer scale, stochastic depth, residual]
//   dd_block_tail_bwd : grad_y (N,HW,C) = scale_b * gamma_c * grad_out (N,C,HW),
//                       grad_gamma_c = sum_{b,p} scale_b * grad_out[b,c,p] * y[b,p,c]   (grad_x = grad_out: no kernel)
// HBM-bound: every element is read once and written once.
#include "dd_common.cuh"

namespace dd {

constexpr int GT = 32;          // tile edge
constexpr int GT_ROWS = 8;      // blockDim.y

// in (B, R, Cc) -> out (B, Cc, R): tile [32 r][32 c] read along c, written along r
__global__ void __launch_bounds__(GT * GT_ROWS) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float t[GT][GT + 1];
  const size_t img = (size_t)blockIdx.z * R * Cc;
  const int r0 = blockIdx.y * GT, c0 = blockIdx.x * GT;
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Cc) ? __ldg(in + img + (size_t)r * Cc + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[img + (size_t)c * R + r] = t[threadIdx.x][i];
  }
}

// out[b][c][p] = x[b][c][p] + scale[b] * gamma[c] * y[b][p][c]
__global__ void __launch_bounds__(GT * GT_ROWS) block_tail_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                      const float* __restrict__ gamma, const float* __restrict__ scale,
                                                                      float* __restrict__ out, int C, int HW) {
  __shared__ float t[GT][GT + 1];
  const int b = blockIdx.z;
  const size_t img = (size_t)b * C * HW;
  const int p0 = blockIdx.x * GT, c0 = blockIdx.y * GT;
  const float sb = scale ? __ldg(scale + b) : 1.f;
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {   // rows = pixels, columns = channels
    const int p = p0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (p < HW && c < C) ? __ldg(y + img + (size_t)p * C + c) * (gamma ? __ldg(gamma + c) : 1.f) * sb : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) {
      const size_t o = img + (size_t)c * HW + p;
      out[o] = __ldg(x + o) + t[threadIdx.x][i];
    }
  }
}

// grad_y[b][p][c] = scale[b] * gamma[c] * g[b][c][p];  grad_gamma[c] += scale[b] * sum_p g[b][c][p] * y[b][p][c]
__global__ void __launch_bounds__(GT * GT_ROWS) block_tail_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                                      const float* __restrict__ gamma, const float* __restrict__ scale,
                                                                      float* __restrict__ grad_y, float* __restrict__ grad_gamma, int C,
                                                                      int HW) {
  __shared__ float t[GT][GT + 1];
  __shared__ float red[GT_ROWS][GT];
  const int b = blockIdx.z;
  const size_t img = (size_t)b * C * HW;
  const int p0 = blockIdx.x * GT, c0 = blockIdx.y * GT;
  const float sb = scale ? __ldg(scale + b) : 1.f;
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {   // rows = channels, columns = pixels
    const int c = c0 + i, p = p0 + threadIdx.x;
    t[i][threadIdx.x] = (p < HW && c < C) ? __ldg(g + img + (size_t)c * HW + p) * sb : 0.f;
  }
  __syncthreads();
  const int c = c0 + threadIdx.x;
  const float gc = (gamma && c < C) ? __ldg(gamma + c) : 1.f;
  float acc = 0.f;
#pragma unroll
  for (int i = threadIdx.y; i < GT; i += GT_ROWS) {
    const int p = p0 + i;
    if (p < HW && c < C) {
      const size_t o = img + (size_t)p * C + c;
      const float gv = t[threadIdx.x][i];
      if (grad_y) grad_y[o] = gv * gc;
      if (grad_gamma) acc = fmaf(gv, __ldg(y + o), acc);
    }
  }
  if (grad_gamma) {   // block-level column sums, one atomic per channel and tile
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < GT_ROWS; ++i) s += red[i][threadIdx.x];
      atomicAdd(grad_gamma + c, s);
    }
  }
}

// ---- 16-byte variants (R % 4 == 0, Cc % 4 == 0, aligned pointers): tile 64 x 64, 256 threads -------------------------------
// Read phase: thread (ty, tx) = (tid / 16, tid % 16) loads the float4 at row ty + 16 i, columns 4 tx .. 4 tx + 3 and stores it
// TRANSPOSED into t[column][row] (pitch 65: the 32 lanes of a warp hit 32 different banks); write phase: the same thread
// reads t[ty + 16 i][4 tx .. 4 tx + 3] (one output row, four consecutive input rows) and stores one float4.
constexpr int VT = 64;
constexpr int VP = VT + 1;

__device__ __forceinline__ void tile_store_transposed(float (*t)[VP], int r, int c4, const float4& v) {
  t[c4][r] = v.x, t[c4 + 1][r] = v.y, t[c4 + 2][r] = v.z, t[c4 + 3][r] = v.w;
}
__device__ __forceinline__ float4 tile_load_row(float (*t)[VP], int c, int r4) { return make_float4(t[c][r4], t[c][r4 + 1], t[c][r4 + 2], t[c][r4 + 3]); }

// in (B, R, Cc) -> out (B, Cc, R)
__global__ void __launch_bounds__(256) transpose4_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float t[VT][VP];
  const size_t img = (size_t)blockIdx.z * R * Cc;
  const int r0 = blockIdx.y * VT, c0 = blockIdx.x * VT;
  const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < R && c0 + tx < Cc) v = __ldg(reinterpret_cast<const float4*>(in + img + (size_t)(r0 + r) * Cc + c0 + tx));
    tile_store_transposed(t, r, tx, v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = ty + 16 * i;
    if (c0 + c < Cc && r0 + tx < R) *reinterpret_cast<float4*>(out + img + (size_t)(c0 + c) * R + r0 + tx) = tile_load_row(t, c, tx);
  }
}

// out[b][c][p] = x[b][c][p] + scale[b] * gamma[c] * y[b][p][c]
__global__ void __launch_bounds__(256) block_tail_fwd4_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gamma,
                                                              const float* __restrict__ scale, float* __restrict__ out, int C, int HW) {
  __shared__ float t[VT][VP];
  const int b = blockIdx.z;
  const size_t img = (size_t)b * C * HW;
  const int p0 = blockIdx.x * VT, c0 = blockIdx.y * VT;
  const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
  const float sb = scale ? __ldg(scale + b) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // rows = pixels, columns = channels
    const int p = ty + 16 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + p < HW && c0 + tx < C) {
      v = __ldg(reinterpret_cast<const float4*>(y + img + (size_t)(p0 + p) * C + c0 + tx));
      const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c0 + tx)) : make_float4(1.f, 1.f, 1.f, 1.f);
      v = make_float4(v.x * gm.x * sb, v.y * gm.y * sb, v.z * gm.z * sb, v.w * gm.w * sb);
    }
    tile_store_transposed(t, p, tx, v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = ty + 16 * i;
    if (c0 + c < C && p0 + tx < HW) {
      const size_t o = img + (size_t)(c0 + c) * HW + p0 + tx;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + o)), tv = tile_load_row(t, c, tx);
      *reinterpret_cast<float4*>(out + o) = make_float4(xv.x + tv.x, xv.y + tv.y, xv.z + tv.z, xv.w + tv.w);
    }
  }
}

// grad_y[b][p][c] = scale[b] * gamma[c] * g[b][c][p];  grad_gamma partial[tile][c] = scale[b] * sum_p g[b][c][p] * y[b][p][c]
__global__ void __launch_bounds__(256) block_tail_bwd4_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ gamma,
                                                              const float* __restrict__ scale, float* __restrict__ grad_y,
                                                              float* __restrict__ grad_gamma, int C, int HW) {
  __shared__ float t[VT][VP];
  __shared__ float red[16][VT];
  const int b = blockIdx.z;
  const size_t img = (size_t)b * C * HW;
  const int p0 = blockIdx.x * VT, c0 = blockIdx.y * VT;
  const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
  const float sb = scale ? __ldg(scale + b) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // rows = channels, columns = pixels
    const int c = ty + 16 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + c < C && p0 + tx < HW) {
      v = __ldg(reinterpret_cast<const float4*>(g + img + (size_t)(c0 + c) * HW + p0 + tx));
      v = make_float4(v.x * sb, v.y * sb, v.z * sb, v.w * sb);
    }
    tile_store_transposed(t, c, tx, v);
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool cok = c0 + tx < C;
  const float4 gm = (gamma && cok) ? __ldg(reinterpret_cast<const float4*>(gamma + c0 + tx)) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ty + 16 * i;
    if (cok && p0 + p < HW) {
      const size_t o = img + (size_t)(p0 + p) * C + c0 + tx;
      const float4 gv = tile_load_row(t, p, tx);
      if (grad_y) *reinterpret_cast<float4*>(grad_y + o) = make_float4(gv.x * gm.x, gv.y * gm.y, gv.z * gm.z, gv.w * gm.w);
      if (grad_gamma) {
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y + o));
        acc.x = fmaf(gv.x, yv.x, acc.x), acc.y = fmaf(gv.y, yv.y, acc.y), acc.z = fmaf(gv.z, yv.z, acc.z), acc.w = fmaf(gv.w, yv.w, acc.w);
      }
    }
  }
  if (grad_gamma) {   // block-level column sums, one atomic per channel and tile
    red[ty][tx] = acc.x, red[ty][tx + 1] = acc.y, red[ty][tx + 2] = acc.z, red[ty][tx + 3] = acc.w;
    __syncthreads();
    if (threadIdx.x < VT && c0 + threadIdx.x < C) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += red[i][threadIdx.x];
      atomicAdd(grad_gamma + c0 + threadIdx.x, s);
    }
  }
}

static bool vec4_ok(int R, int Cc, const void* a, const void* b, const void* c = nullptr, const void* d = nullptr) {
  return R % 4 == 0 && Cc % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

static int check_dims(const char* what, int B, int C, int HW) {
  DD_REQUIRE(B > 0 && C > 0 && HW > 0, "%s: bad shape B=%d C=%d HW=%d", what, B, C, HW);
  DD_REQUIRE(B <= 65535 && (C + GT - 1) / GT <= 65535, "%s: batch / channel count too large for one launch", what);
  return DD_OK;
}

}  // namespace dd

extern "C" {

int dd_nchw_to_nhwc(const float* x, int B, int C, int HW, float* out, void* stream) {
  DD_REQUIRE(x && out, "dd_nchw_to_nhwc: NULL pointer");
  if (int rc = dd::check_dims("dd_nchw_to_nhwc", B, C, HW)) return rc;
  if (dd::vec4_ok(C, HW, x, out)) {
    dim3 grid4((HW + dd::VT - 1) / dd::VT, (C + dd::VT - 1) / dd::VT, B);
    dd::transpose4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(x, out, C, HW);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  dim3 grid((HW + dd::GT - 1) / dd::GT, (C + dd::GT - 1) / dd::GT, B);   // in (B, R = C, Cc = HW)
  dd::transpose_kernel<<<grid, dim3(dd::GT, dd::GT_ROWS), 0, (cudaStream_t)stream>>>(x, out, C, HW);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_nhwc_to_nchw(const float* x, int B, int C, int HW, float* out, void* stream) {
  DD_REQUIRE(x && out, "dd_nhwc_to_nchw: NULL pointer");
  if (int rc = dd::check_dims("dd_nhwc_to_nchw", B, C, HW)) return rc;
  dim3 grid((C + dd::GT - 1) / dd::GT, (HW + dd::GT - 1) / dd::GT, B);   // in (B, R = HW, Cc = C)
  DD_REQUIRE(grid.y <= 65535, "dd_nhwc_to_nchw: HW too large");
  if (dd::vec4_ok(HW, C, x, out)) {
    dim3 grid4((C + dd::VT - 1) / dd::VT, (HW + dd::VT - 1) / dd::VT, B);
    dd::transpose4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(x, out, HW, C);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  dd::transpose_kernel<<<grid, dim3(dd::GT, dd::GT_ROWS), 0, (cudaStream_t)stream>>>(x, out, HW, C);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_block_tail_fwd(const float* x, const float* y, const float* gamma, const float* scale, int B, int C, int HW, float* out,
                      void* stream) {
  DD_REQUIRE(x && y && out, "dd_block_tail_fwd: NULL pointer");
  if (int rc = dd::check_dims("dd_block_tail_fwd", B, C, HW)) return rc;
  if (dd::vec4_ok(C, HW, x, y, out, gamma)) {
    dim3 grid4((HW + dd::VT - 1) / dd::VT, (C + dd::VT - 1) / dd::VT, B);
    dd::block_tail_fwd4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(x, y, gamma, scale, out, C, HW);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  dim3 grid((HW + dd::GT - 1) / dd::GT, (C + dd::GT - 1) / dd::GT, B);
  dd::block_tail_fwd_kernel<<<grid, dim3(dd::GT, dd::GT_ROWS), 0, (cudaStream_t)stream>>>(x, y, gamma, scale, out, C, HW);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_block_tail_bwd(const float* grad_out, const float* y, const float* gamma, const float* scale, int B, int C, int HW,
                      float* grad_y, float* grad_gamma, void* stream) {
  DD_REQUIRE(grad_out && (grad_y || grad_gamma), "dd_block_tail_bwd: NULL pointer");
  DD_REQUIRE(!grad_gamma || y, "dd_block_tail_bwd: y is needed for grad_gamma");
  if (int rc = dd::check_dims("dd_block_tail_bwd", B, C, HW)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_gamma) DD_CHECK_CUDA(cudaMemsetAsync(grad_gamma, 0, (size_t)C * sizeof(float), st));
  if (dd::vec4_ok(C, HW, grad_out, y, grad_y, gamma)) {
    dim3 grid4((HW + dd::VT - 1) / dd::VT, (C + dd::VT - 1) / dd::VT, B);
    dd::block_tail_bwd4_kernel<<<grid4, 256, 0, st>>>(grad_out, y, gamma, scale, grad_y, grad_gamma, C, HW);
    dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
    return DD_OK;
  }
  dim3 grid((HW + dd::GT - 1) / dd::GT, (C + dd::GT - 1) / dd::GT, B);
  dd::block_tail_bwd_kernel<<<grid, dim3(dd::GT, dd::GT_ROWS), 0, st>>>(grad_out, y, gamma, scale, grad_y, grad_gamma, C, HW);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
