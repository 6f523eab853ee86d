// Pose head epilogue: spatial mean (x 0.01) and axis-angle/translation -> 4x4 transform, forward and backward
// (pose_decoder.py:39-44, networks/layers.py:7-82).  Replaces ~60 tiny ATen launches per call by two.
#include "dd_common.cuh"

namespace dd {

__global__ void pose_mean_fwd_kernel(const float* __restrict__ x, int BC, int hw, float scale, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per (b, c)
  if (row >= BC) return;
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < hw; i += 32) acc += __ldg(x + (size_t)row * hw + i);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) out[row] = scale * (acc / (float)hw);
}

__global__ void pose_mean_bwd_kernel(const float* __restrict__ go, int BC, int hw, float scale, float* __restrict__ gx) {
  const size_t n = (size_t)BC * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    gx[i] = __ldg(go + i / hw) * (scale / (float)hw);
}

struct Rod {
  float theta, inv, x, y, z, ca, sa, C;
};

__device__ __forceinline__ Rod rodrigues_terms(const float* v) {
  Rod r;
  r.theta = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);   // torch.norm(vec, 2, 2, True)
  r.inv = 1.f / (r.theta + 1e-7f);
  r.x = v[0] * r.inv, r.y = v[1] * r.inv, r.z = v[2] * r.inv;
  r.ca = cosf(r.theta), r.sa = sinf(r.theta);
  r.C = 1.f - r.ca;
  return r;
}

__device__ __forceinline__ void rotation(const Rod& r, float R[3][3]) {
  const float xs = r.x * r.sa, ys = r.y * r.sa, zs = r.z * r.sa;
  const float xC = r.x * r.C, yC = r.y * r.C, zC = r.z * r.C;
  const float xyC = r.x * yC, yzC = r.y * zC, zxC = r.z * xC;
  R[0][0] = r.x * xC + r.ca, R[0][1] = xyC - zs, R[0][2] = zxC + ys;
  R[1][0] = xyC + zs, R[1][1] = r.y * yC + r.ca, R[1][2] = yzC - xs;
  R[2][0] = zxC - ys, R[2][1] = yzC + xs, R[2][2] = r.z * zC + r.ca;
}

__global__ void pose_matrix_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int B, int invert,
                                       float* __restrict__ T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float v[3] = {aa[b * 3], aa[b * 3 + 1], aa[b * 3 + 2]};
  const float t[3] = {tr[b * 3], tr[b * 3 + 1], tr[b * 3 + 2]};
  float R[3][3];
  rotation(rodrigues_terms(v), R);
  float* M = T + b * 16;
  if (invert) {   // R^T @ Trans(-t)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[j][i];
      M[i * 4 + 3] = R[0][i] * -t[0] + R[1][i] * -t[1] + R[2][i] * -t[2];
    }
  } else {        // Trans(t) @ R
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i][j];
      M[i * 4 + 3] = t[i];
    }
  }
  M[12] = 0.f, M[13] = 0.f, M[14] = 0.f, M[15] = 1.f;
}

__global__ void pose_matrix_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ gT,
                                       int B, int invert, float* __restrict__ gaa, float* __restrict__ gtr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float v[3] = {aa[b * 3], aa[b * 3 + 1], aa[b * 3 + 2]};
  const float t[3] = {tr[b * 3], tr[b * 3 + 1], tr[b * 3 + 2]};
  const Rod r = rodrigues_terms(v);
  float R[3][3];
  rotation(r, R);
  const float* g = gT + b * 16;
  float gR[3][3], gt[3];
  if (invert) {   // M[:3,:3] = R^T, M[:3,3] = R^T t' with t' = -t
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) gR[j][i] = g[i * 4 + j] + g[i * 4 + 3] * -t[j];   // d/dR'[i][j], R' = R^T
#pragma unroll
    for (int j = 0; j < 3; ++j) gt[j] = -(R[j][0] * g[3] + R[j][1] * g[7] + R[j][2] * g[11]);   // -(R'^T gM[:3,3])
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) gR[i][j] = g[i * 4 + j];
      gt[i] = g[i * 4 + 3];
    }
  }
  const float x = r.x, y = r.y, z = r.z, sa = r.sa, ca = r.ca, C = r.C;
  float gx = 0.f, gy = 0.f, gz = 0.f, gC = 0.f, gca = 0.f, gsa = 0.f;
  gx += gR[0][0] * 2.f * x * C, gC += gR[0][0] * x * x, gca += gR[0][0];
  gy += gR[1][1] * 2.f * y * C, gC += gR[1][1] * y * y, gca += gR[1][1];
  gz += gR[2][2] * 2.f * z * C, gC += gR[2][2] * z * z, gca += gR[2][2];
  // xyC -+ z sa
  gx += (gR[0][1] + gR[1][0]) * y * C, gy += (gR[0][1] + gR[1][0]) * x * C, gC += (gR[0][1] + gR[1][0]) * x * y;
  gz += (gR[1][0] - gR[0][1]) * sa, gsa += (gR[1][0] - gR[0][1]) * z;
  // zxC +- y sa
  gz += (gR[0][2] + gR[2][0]) * x * C, gx += (gR[0][2] + gR[2][0]) * z * C, gC += (gR[0][2] + gR[2][0]) * z * x;
  gy += (gR[0][2] - gR[2][0]) * sa, gsa += (gR[0][2] - gR[2][0]) * y;
  // yzC -+ x sa
  gy += (gR[1][2] + gR[2][1]) * z * C, gz += (gR[1][2] + gR[2][1]) * y * C, gC += (gR[1][2] + gR[2][1]) * y * z;
  gx += (gR[2][1] - gR[1][2]) * sa, gsa += (gR[2][1] - gR[1][2]) * x;
  gca -= gC;                                     // C = 1 - cos
  float gtheta = -sa * gca + ca * gsa;           // cos, sin
  const float ga[3] = {gx, gy, gz};
  float gv[3];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gv[i] = ga[i] * r.inv;                       // axis = v / (theta + eps)
    dot += ga[i] * v[i];
  }
  gtheta += -dot * r.inv * r.inv;
  if (r.theta > 0.f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) gv[i] += gtheta * v[i] / r.theta;   // theta = |v|
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) gaa[b * 3 + i] = gv[i], gtr[b * 3 + i] = gt[i];
}

}  // namespace dd

extern "C" {
using namespace dd;

int dd_pose_mean_fwd(const float* x, int BC, int hw, float scale, float* out, void* stream) {
  DD_REQUIRE(x && out && BC > 0 && hw > 0, "dd_pose_mean_fwd: bad arguments");
  pose_mean_fwd_kernel<<<(BC + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, BC, hw, scale, out); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_pose_mean_bwd(const float* grad_out, int BC, int hw, float scale, float* grad_x, void* stream) {
  DD_REQUIRE(grad_out && grad_x && BC > 0 && hw > 0, "dd_pose_mean_bwd: bad arguments");
  const size_t n = (size_t)BC * hw;
  pose_mean_bwd_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(grad_out, BC, hw, scale, grad_x);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_pose_matrix_fwd(const float* axisangle, const float* translation, int B, int invert, float* T, void* stream) {
  DD_REQUIRE(axisangle && translation && T && B > 0, "dd_pose_matrix_fwd: bad arguments");
  pose_matrix_fwd_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(axisangle, translation, B, invert, T); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_T, int B, int invert,
                       float* grad_axisangle, float* grad_translation, void* stream) {
  DD_REQUIRE(axisangle && translation && grad_T && grad_axisangle && grad_translation && B > 0, "dd_pose_matrix_bwd: bad arguments");
  pose_matrix_bwd_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(axisangle, translation, grad_T, B, invert, grad_axisangle,
                                                                         grad_translation);
  dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
