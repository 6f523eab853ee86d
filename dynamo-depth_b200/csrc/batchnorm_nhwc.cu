// Training-mode BatchNorm2d fused with its activation and residual add on channels_last (N,H,W,C) activations:
//   y = act( (x - mean_c) * invstd_c * gamma_c + beta_c  [+ residual] ),   act in {none, ReLU, exact GELU}
// -- the BasicBlock pattern of the ResNet trunks (torchvision BasicBlock.forward as used by the reference's
// networks/resnet_encoder.py:16-20,:125-134: conv-bn-relu, conv-bn-(+identity)-relu, stem conv1-bn1-relu) and the Lite-Mono
// stem's BNGELU (networks/depth_encoder.py:113-122) when that stem runs channels_last (no NCHW<->NHWC conversions around its
// cuDNN convolutions).  PyTorch runs BN, the add and the ReLU as separate passes over the activation (plus their backward
// passes); here each direction is two streaming passes over rows of C channels:
//   forward   stats (shifted sums per channel) -> finalize (mean / invstd / running statistics) -> apply (+res, act)
//   backward  g = grad_y * act'(.) (ReLU: mask from the saved output; GELU: re-derived from x), sums of g and g * xhat
//             -> finalize -> grad_x = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)), grad_residual = g
// Thread = one 16-byte channel quad of a strip of rows (256 threads = C/4 quads x 1024/C row groups): a warp reads whole
// rows; per-CTA partials and a fixed-order finalize (deterministic, no atomics).
#include "dd_common.cuh"

namespace dd {

constexpr int BNH_THREADS = 256;
constexpr int BNH_MAX_CHUNKS = 592;

enum { BNH_NONE = 0, BNH_RELU = 1, BNH_GELU = 2 };

struct BnhArgs {
  const float* x;
  const float* y_saved;   // backward, ReLU: the forward output (mask)
  const float* gy;
  const float* residual;  // forward
  const float* gamma;
  const float* beta;
  float* y;               // forward: y; backward: grad_x
  float* gres;            // backward: grad_residual (= g), or NULL
  float* mean;            // (C)
  float* invstd;          // (C)
  float* running_mean;
  float* running_var;
  float* grad_gamma;
  float* grad_beta;
  float* partial;         // [chunks][2][C]
  float* coef;            // backward finalize -> [2][C]: mean(g), mean(g * xhat)
  long long M;            // rows = N*H*W
  int C, chunks, rows_per_cta;
  float eps, momentum;
  int act;
};

__device__ __forceinline__ float bnh_gelu(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float bnh_gelu_grad(float z) {
  const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
  return fmaf(z, 0.3989422804014327f * __expf(-0.5f * z * z), cdf);
}

// column sums of two float4 accumulators over the row groups of the CTA -> partial[chunk][0 | 1][C]
__device__ __forceinline__ void bnh_reduce_store(const BnhArgs& a, const float4& s1, const float4& s2, float* red /* [2][RG][C] */) {
  const int Q = a.C >> 2, RG = BNH_THREADS / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  reinterpret_cast<float4*>(red)[rg * Q + q] = s1;
  reinterpret_cast<float4*>(red)[(RG + rg) * Q + q] = s2;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * a.C; i += BNH_THREADS) {
    const int which = i / a.C, c = i - which * a.C;
    float s = 0.f;
    for (int g = 0; g < RG; ++g) s += red[(which * RG + g) * a.C + c];
    a.partial[((size_t)blockIdx.x * 2 + which) * a.C + c] = s;
  }
}

__global__ void __launch_bounds__(BNH_THREADS) bnh_stats_kernel(const __grid_constant__ BnhArgs a) {
  extern __shared__ float red[];
  const int Q = a.C >> 2, RG = BNH_THREADS / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta, r1 = min(a.M, r0 + a.rows_per_cta);
  const float4* xp = reinterpret_cast<const float4*>(a.x);
  const float4 k0 = __ldg(xp + q);   // shift: first row of the tensor
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  long long r = r0 + rg;
  for (; r + 3ll * RG < r1; r += 4ll * RG) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(xp + (r + (long long)j * RG) * Q + q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d0 = v[j].x - k0.x, d1 = v[j].y - k0.y, d2 = v[j].z - k0.z, d3 = v[j].w - k0.w;
      s1.x += d0, s1.y += d1, s1.z += d2, s1.w += d3;
      s2.x = fmaf(d0, d0, s2.x), s2.y = fmaf(d1, d1, s2.y), s2.z = fmaf(d2, d2, s2.z), s2.w = fmaf(d3, d3, s2.w);
    }
  }
  for (; r < r1; r += RG) {
    const float4 v = __ldg(xp + r * Q + q);
    const float d0 = v.x - k0.x, d1 = v.y - k0.y, d2 = v.z - k0.z, d3 = v.w - k0.w;
    s1.x += d0, s1.y += d1, s1.z += d2, s1.w += d3;
    s2.x = fmaf(d0, d0, s2.x), s2.y = fmaf(d1, d1, s2.y), s2.z = fmaf(d2, d2, s2.z), s2.w = fmaf(d3, d3, s2.w);
  }
  bnh_reduce_store(a, s1, s2, red);
}

// partial[chunk][which][c] summed over the chunks: CTA = 32 channels x 32 chunk groups (1024 threads), four independent loads per
// step -- a first version with 8 groups and one load per step needed 74 dependent L2 round trips = 60 us for a 2-CTA kernel (ncu)
constexpr int BNH_FIN_GROUPS = 32;
__device__ __forceinline__ bool bnh_sum_partials(const BnhArgs& a, int c, double& S1, double& S2) {
  __shared__ double sh[2][BNH_FIN_GROUPS][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  S1 = 0.0, S2 = 0.0;
  if (c < a.C) {
    for (int k = grp; k < a.chunks; k += 4 * BNH_FIN_GROUPS) {
      float v1[4], v2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j * BNH_FIN_GROUPS;
        const bool ok = kk < a.chunks;
        v1[j] = ok ? a.partial[((size_t)kk * 2) * a.C + c] : 0.f;
        v2[j] = ok ? a.partial[((size_t)kk * 2 + 1) * a.C + c] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) S1 += (double)v1[j], S2 += (double)v2[j];
    }
  }
  sh[0][grp][lane] = S1, sh[1][grp][lane] = S2;
  __syncthreads();
  if (grp != 0 || c >= a.C) return false;
#pragma unroll 8
  for (int g = 1; g < BNH_FIN_GROUPS; ++g) S1 += sh[0][g][lane], S2 += sh[1][g][lane];
  return true;
}

// partials -> mean / invstd (+ running statistics); grid = ceil(C / 32), 1024 threads
__global__ void __launch_bounds__(32 * BNH_FIN_GROUPS) bnh_finalize_kernel(const __grid_constant__ BnhArgs a) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double S1, S2;
  if (!bnh_sum_partials(a, c, S1, S2)) return;
  const double n = (double)a.M, m1 = S1 / n;
  double var = S2 / n - m1 * m1;
  var = var < 0.0 ? 0.0 : var;
  const float mean = (float)((double)__ldg(a.x + c) + m1);
  a.mean[c] = mean;
  a.invstd[c] = (float)(1.0 / sqrt(var + (double)a.eps));
  if (a.running_mean) a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
  if (a.running_var) a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)(var * (n / (n > 1.0 ? n - 1.0 : 1.0)));
}

__global__ void __launch_bounds__(BNH_THREADS) bnh_apply_kernel(const __grid_constant__ BnhArgs a) {
  const int Q = a.C >> 2, RG = BNH_THREADS / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta, r1 = min(a.M, r0 + a.rows_per_cta);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + q), is = __ldg(reinterpret_cast<const float4*>(a.invstd) + q);
  const float4 gm = a.gamma ? __ldg(reinterpret_cast<const float4*>(a.gamma) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = a.beta ? __ldg(reinterpret_cast<const float4*>(a.beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 sc = make_float4(is.x * gm.x, is.y * gm.y, is.z * gm.z, is.w * gm.w);
  const float4* xp = reinterpret_cast<const float4*>(a.x);
  const float4* rp = reinterpret_cast<const float4*>(a.residual);
  float4* yp = reinterpret_cast<float4*>(a.y);
  auto one = [&](float4 v, float4 res) {
    float4 z = make_float4(fmaf(v.x - mu.x, sc.x, bt.x) + res.x, fmaf(v.y - mu.y, sc.y, bt.y) + res.y, fmaf(v.z - mu.z, sc.z, bt.z) + res.z,
                           fmaf(v.w - mu.w, sc.w, bt.w) + res.w);
    if (a.act == BNH_RELU) z = make_float4(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f));
    else if (a.act == BNH_GELU) z = make_float4(bnh_gelu(z.x), bnh_gelu(z.y), bnh_gelu(z.z), bnh_gelu(z.w));
    return z;
  };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  long long r = r0 + rg;
  for (; r + 3ll * RG < r1; r += 4ll * RG) {
    float4 v[4], res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long o = (r + (long long)j * RG) * Q + q;
      v[j] = __ldg(xp + o);
      res[j] = rp ? __ldg(rp + o) : zero;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) yp[(r + (long long)j * RG) * Q + q] = one(v[j], res[j]);
  }
  for (; r < r1; r += RG) yp[r * Q + q] = one(__ldg(xp + r * Q + q), rp ? __ldg(rp + r * Q + q) : zero);
}

// g = grad_y * act'(.) for one quad
__device__ __forceinline__ float4 bnh_g(const BnhArgs& a, const float4& gy, const float4& xh, const float4& ys, const float4& gm, const float4& bt) {
  if (a.act == BNH_RELU) return make_float4(ys.x > 0.f ? gy.x : 0.f, ys.y > 0.f ? gy.y : 0.f, ys.z > 0.f ? gy.z : 0.f, ys.w > 0.f ? gy.w : 0.f);
  if (a.act == BNH_GELU)
    return make_float4(gy.x * bnh_gelu_grad(fmaf(xh.x, gm.x, bt.x)), gy.y * bnh_gelu_grad(fmaf(xh.y, gm.y, bt.y)),
                       gy.z * bnh_gelu_grad(fmaf(xh.z, gm.z, bt.z)), gy.w * bnh_gelu_grad(fmaf(xh.w, gm.w, bt.w)));
  return gy;
}

__global__ void __launch_bounds__(BNH_THREADS) bnh_bwd_stats_kernel(const __grid_constant__ BnhArgs a) {
  extern __shared__ float red[];
  const int Q = a.C >> 2, RG = BNH_THREADS / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta, r1 = min(a.M, r0 + a.rows_per_cta);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + q), is = __ldg(reinterpret_cast<const float4*>(a.invstd) + q);
  const float4 gm = a.gamma ? __ldg(reinterpret_cast<const float4*>(a.gamma) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = a.beta ? __ldg(reinterpret_cast<const float4*>(a.beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* xp = reinterpret_cast<const float4*>(a.x);
  const float4* gp = reinterpret_cast<const float4*>(a.gy);
  const float4* yp = reinterpret_cast<const float4*>(a.y_saved);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const bool relu = a.act == BNH_RELU;
  for (long long r = r0 + rg; r < r1; r += 2ll * RG) {
    float4 v[2], gy[2], ys[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long rr = r + (long long)j * RG;
      ok[j] = rr < r1;
      const long long o = (ok[j] ? rr : r) * Q + q;
      v[j] = __ldg(xp + o), gy[j] = __ldg(gp + o);
      ys[j] = relu ? __ldg(yp + o) : v[j];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (!ok[j]) continue;
      const float4 xh = make_float4((v[j].x - mu.x) * is.x, (v[j].y - mu.y) * is.y, (v[j].z - mu.z) * is.z, (v[j].w - mu.w) * is.w);
      const float4 g = bnh_g(a, gy[j], xh, ys[j], gm, bt);
      s1.x += g.x, s1.y += g.y, s1.z += g.z, s1.w += g.w;
      s2.x = fmaf(g.x, xh.x, s2.x), s2.y = fmaf(g.y, xh.y, s2.y), s2.z = fmaf(g.z, xh.z, s2.z), s2.w = fmaf(g.w, xh.w, s2.w);
    }
  }
  bnh_reduce_store(a, s1, s2, red);
}

__global__ void __launch_bounds__(32 * BNH_FIN_GROUPS) bnh_bwd_finalize_kernel(const __grid_constant__ BnhArgs a) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double S1, S2;
  if (!bnh_sum_partials(a, c, S1, S2)) return;
  if (a.grad_beta) a.grad_beta[c] = (float)S1;
  if (a.grad_gamma) a.grad_gamma[c] = (float)S2;
  a.coef[c] = (float)(S1 / (double)a.M);
  a.coef[a.C + c] = (float)(S2 / (double)a.M);
}

__global__ void __launch_bounds__(BNH_THREADS) bnh_bwd_apply_kernel(const __grid_constant__ BnhArgs a) {
  const int Q = a.C >> 2, RG = BNH_THREADS / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta, r1 = min(a.M, r0 + a.rows_per_cta);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(a.mean) + q), is = __ldg(reinterpret_cast<const float4*>(a.invstd) + q);
  const float4 gm = a.gamma ? __ldg(reinterpret_cast<const float4*>(a.gamma) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = a.beta ? __ldg(reinterpret_cast<const float4*>(a.beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mg = __ldg(reinterpret_cast<const float4*>(a.coef) + q), mgx = __ldg(reinterpret_cast<const float4*>(a.coef + a.C) + q);
  const float4 sc = make_float4(is.x * gm.x, is.y * gm.y, is.z * gm.z, is.w * gm.w);
  const float4* xp = reinterpret_cast<const float4*>(a.x);
  const float4* gp = reinterpret_cast<const float4*>(a.gy);
  const float4* yp = reinterpret_cast<const float4*>(a.y_saved);
  float4* dxp = reinterpret_cast<float4*>(a.y);
  float4* drp = reinterpret_cast<float4*>(a.gres);
  const bool relu = a.act == BNH_RELU;
  for (long long r = r0 + rg; r < r1; r += 2ll * RG) {
    float4 v[2], gy[2], ys[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long rr = r + (long long)j * RG;
      ok[j] = rr < r1;
      const long long o = (ok[j] ? rr : r) * Q + q;
      v[j] = __ldg(xp + o), gy[j] = __ldg(gp + o);
      ys[j] = relu ? __ldg(yp + o) : v[j];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (!ok[j]) continue;
      const long long o = (r + (long long)j * RG) * Q + q;
      const float4 xh = make_float4((v[j].x - mu.x) * is.x, (v[j].y - mu.y) * is.y, (v[j].z - mu.z) * is.z, (v[j].w - mu.w) * is.w);
      const float4 g = bnh_g(a, gy[j], xh, ys[j], gm, bt);
      if (drp) drp[o] = g;
      if (dxp)
        dxp[o] = make_float4(sc.x * ((g.x - mg.x) - xh.x * mgx.x), sc.y * ((g.y - mg.y) - xh.y * mgx.y), sc.z * ((g.z - mg.z) - xh.z * mgx.z),
                             sc.w * ((g.w - mg.w) - xh.w * mgx.w));
    }
  }
}

static int bnh_plan(BnhArgs& a, const char* what) {
  DD_REQUIRE(a.M > 0 && a.C > 0, "%s: bad shape M=%lld C=%d", what, a.M, a.C);
  const int Q = a.C / 4;
  DD_REQUIRE(a.C % 4 == 0 && Q >= 1 && Q <= BNH_THREADS && BNH_THREADS % Q == 0, "%s: C must be 4 * a divisor of 256 (got %d)", what, a.C);
  const int RG = BNH_THREADS / Q;
  long long chunks = (a.M + (long long)RG * 8 - 1) / ((long long)RG * 8);   // >= 8 rows per thread
  chunks = chunks > BNH_MAX_CHUNKS ? BNH_MAX_CHUNKS : (chunks < 1 ? 1 : chunks);
  a.rows_per_cta = (int)((a.M + chunks - 1) / chunks);
  a.chunks = (int)((a.M + a.rows_per_cta - 1) / a.rows_per_cta);
  return DD_OK;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dd

extern "C" {

size_t dd_bn_nhwc_workspace_bytes(int C) { return C > 0 ? ((size_t)dd::BNH_MAX_CHUNKS * 2 + 2) * C * sizeof(float) : 0; }

int dd_bn_act_nhwc_fwd(const float* x, const float* residual, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       int act, float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* workspace,
                       size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && y && save_mean && save_invstd, "dd_bn_act_nhwc_fwd: NULL pointer");
  DD_REQUIRE(act >= 0 && act <= 2, "dd_bn_act_nhwc_fwd: bad activation %d", act);
  DD_REQUIRE(!(act == BNH_GELU && residual != nullptr), "dd_bn_act_nhwc_fwd: GELU with a residual is not built (its derivative is re-derived from x)");
  DD_REQUIRE(al16(x) && al16(y) && al16(residual) && al16(gamma) && al16(beta) && al16(save_mean) && al16(save_invstd),
             "dd_bn_act_nhwc_fwd: pointers must be 16-byte aligned");
  BnhArgs a = {};
  a.x = x, a.residual = residual, a.gamma = gamma, a.beta = beta, a.y = y, a.mean = save_mean, a.invstd = save_invstd;
  a.running_mean = running_mean, a.running_var = running_var, a.M = M, a.C = C, a.eps = eps, a.momentum = momentum, a.act = act;
  if (int rc = bnh_plan(a, "dd_bn_act_nhwc_fwd")) return rc;
  if (!workspace || workspace_bytes < dd_bn_nhwc_workspace_bytes(C)) {
    set_error("dd_bn_act_nhwc_fwd: workspace too small (%zu < %zu)", workspace_bytes, dd_bn_nhwc_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  a.partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * (BNH_THREADS / (C / 4)) * C * sizeof(float);   // 2 x RG x C floats = 8 KB
  bnh_stats_kernel<<<a.chunks, BNH_THREADS, smem, st>>>(a);
  bnh_finalize_kernel<<<(C + 31) / 32, 32 * BNH_FIN_GROUPS, 0, st>>>(a);
  bnh_apply_kernel<<<a.chunks, BNH_THREADS, 0, st>>>(a);
  count_launches(3);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int dd_bn_act_nhwc_bwd(const float* x, const float* y, const float* grad_y, long long M, int C, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, int act, float* grad_x, float* grad_residual, float* grad_gamma,
                       float* grad_beta, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dd;
  DD_REQUIRE(x && grad_y && save_mean && save_invstd, "dd_bn_act_nhwc_bwd: NULL pointer");
  DD_REQUIRE(act >= 0 && act <= 2, "dd_bn_act_nhwc_bwd: bad activation %d", act);
  DD_REQUIRE(act != BNH_RELU || y != nullptr, "dd_bn_act_nhwc_bwd: the forward output is needed for the ReLU mask");
  DD_REQUIRE(al16(x) && al16(y) && al16(grad_y) && al16(grad_x) && al16(grad_residual) && al16(gamma) && al16(beta) && al16(save_mean) &&
                 al16(save_invstd), "dd_bn_act_nhwc_bwd: pointers must be 16-byte aligned");
  BnhArgs a = {};
  a.x = x, a.y_saved = y, a.gy = grad_y, a.gamma = gamma, a.beta = beta, a.y = grad_x, a.gres = grad_residual;
  a.mean = const_cast<float*>(save_mean), a.invstd = const_cast<float*>(save_invstd), a.grad_gamma = grad_gamma, a.grad_beta = grad_beta;
  a.M = M, a.C = C, a.act = act;
  if (int rc = bnh_plan(a, "dd_bn_act_nhwc_bwd")) return rc;
  if (!workspace || workspace_bytes < dd_bn_nhwc_workspace_bytes(C)) {
    set_error("dd_bn_act_nhwc_bwd: workspace too small (%zu < %zu)", workspace_bytes, dd_bn_nhwc_workspace_bytes(C));
    return DD_ERR_WORKSPACE;
  }
  a.partial = reinterpret_cast<float*>(workspace);
  a.coef = a.partial + (size_t)BNH_MAX_CHUNKS * 2 * C;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * (BNH_THREADS / (C / 4)) * C * sizeof(float);
  bnh_bwd_stats_kernel<<<a.chunks, BNH_THREADS, smem, st>>>(a);
  bnh_bwd_finalize_kernel<<<(C + 31) / 32, 32 * BNH_FIN_GROUPS, 0, st>>>(a);
  count_launches(2);
  if (grad_x || grad_residual) {
    bnh_bwd_apply_kernel<<<a.chunks, BNH_THREADS, 0, st>>>(a);
    count_launches(1);
  }
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // extern "C"
