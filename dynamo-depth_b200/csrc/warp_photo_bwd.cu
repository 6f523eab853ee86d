// Fused view synthesis + photometric loss, backward (dd_warp_photo_bwd).
//
// One CTA owns a 32x16 tile of one image and, per pyramid level,
//   stage A  stages the warped source frames over the tile + 2-pixel halo (36x20) in shared memory:
//            read back from the forward pass' warped images when they were kept (SAVED, 24 B per
//            pixel and level instead of ~600 instructions of geometry + 24 gathers), else re-warped;
//   stage B  re-derives the 3x3 SSIM statistics with a register ring marching down each column of
//            the tile + 1-pixel halo (34x18), the candidate losses and the per-pixel argmin, and
//            stores, for the selected frame only, the three coefficient maps of
//            d(loss)/d(mu_x, E[x^2], E[xy]);
//   stage C  box-sums the coefficient maps (ReflectionPad2d fold-back = weight 2 on the rows /
//            columns next to the border), adds the L1 term, re-gathers the four bilinear taps and
//            chains through projection, pose, (scene flow, motion mask,) back-projection and
//            disp->depth (appendix A.3); pose gradients are reduced per CTA;
//   stage D  transposes the bilinear up-sampling of disp_s / flow_s / mask_s separably (rows, then
//            columns) inside shared memory and flushes the low-resolution tile (plain stores at
//            level 0, atomics above).
#include "warp_photo.cuh"

namespace dd {

constexpr int BT_W = 32, BT_H = 16;           // backward tile
constexpr int H2_W = BT_W + 4, H2_H = BT_H + 4;   // 36 x 20 (halo 2)
constexpr int H1_W = BT_W + 2, H1_H = BT_H + 2;   // 34 x 18 (halo 1)
// odd pitches: conflict-free rows, and two CTAs (dynamic + static + 1 KB each) stay inside the 164 KB carve-out step
constexpr int PITCH2 = 37;
constexpr int PLANE2 = H2_H * PITCH2;          // 740
constexpr int CPITCH = 35;
constexpr int CPLANE = H1_H * CPITCH;          // 630
constexpr int BP_PLANE = (BT_H / 2 + 2) * (BT_W / 2 + 2);   // 180: low-resolution patch plane (level 1 is the largest)
constexpr int GPLANE = BT_W * BT_H;            // 512

struct BwdArgs {
  dd_warp_desc d;
  dd_warp_grads g;
  const float* resid_saved[DD_MAX_SCALES][DD_MAX_FRAMES];
  const float* warped_saved[DD_MAX_SCALES][DD_MAX_FRAMES];   // (B,3,H,W) from the forward pass (SAVED kernels)
  const float* grad_sums;
  float* partial_T;   // [num_ctas][2][12]
  float min_disp, disp_range;
  cudaTextureObject_t tex[DD_MAX_FRAMES];   // source frames as textures (0: gather with plain loads)
};

// smem (floats): Y[3] | X[2][3] (PLANE2 each) | LID[2][CPLANE] | COEF[10][CPLANE] | GT[9][GPLANE]
constexpr int SM_Y = 0;
constexpr int SM_X = 3 * PLANE2;
constexpr int SM_LID = 9 * PLANE2;
constexpr int SM_COEF = SM_LID + 2 * CPLANE;
constexpr int SM_GT = SM_COEF + 10 * CPLANE;
constexpr int SM_PATCH = SM_GT + 9 * GPLANE;   // [9][BP_PLANE]: disp | flow f0 xyz | flow f1 xyz | mask f0 | mask f1
constexpr int SM_TOTAL = SM_PATCH + 9 * BP_PLANE;

// Per-position SSIM statistics of all three channels and both frames from 3x3 windows, produced by a
// register ring that marches down one column of the halo-2 tiles (3 horizontal taps per row from
// shared memory).  `emit(k, st)` is called for output row k (0..R-1) of the run.
struct PosStats {
  float mu_y[3], sig_y[3];
  float mu_x[2][3], sig_x[2][3], sig_xy[2][3];
  float yc[3], xc[2][3];   // centre values (L1 term)
};

template <int F, int R, typename Emit>
__device__ __forceinline__ void march_stats(const float* __restrict__ Y, const float* __restrict__ X, int row0, int col0,
                                            Emit emit) {
  const float inv9 = 1.f / 9.f;
  float ay[3], by[3], ayy[3], byy[3];
  float ax[2][3], bx[2][3], axx[2][3], bxx[2][3], axy[2][3], bxy[2][3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    ay[ch] = by[ch] = ayy[ch] = byy[ch] = 0.f;
#pragma unroll
    for (int f = 0; f < 2; ++f) ax[f][ch] = bx[f][ch] = axx[f][ch] = bxx[f][ch] = axy[f][ch] = bxy[f][ch] = 0.f;
  }
#pragma unroll
  for (int rr = 0; rr < R + 2; ++rr) {
    const int o = (row0 + rr) * PITCH2 + col0;
    PosStats st;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* Yc = Y + ch * PLANE2 + o;
      const float yl = Yc[0], ym = Yc[1], yr = Yc[2];
      const float hy = yl + ym + yr, hyy = yl * yl + ym * ym + yr * yr;
      if (rr >= 2) {
        st.mu_y[ch] = (ay[ch] + hy) * inv9;
        st.sig_y[ch] = (ayy[ch] + hyy) * inv9 - st.mu_y[ch] * st.mu_y[ch];
        st.yc[ch] = Yc[1 - PITCH2];
      }
      ay[ch] = by[ch] + hy, by[ch] = hy, ayy[ch] = byy[ch] + hyy, byy[ch] = hyy;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float* Xc = X + (f * 3 + ch) * PLANE2 + o;
        const float xl = Xc[0], xm = Xc[1], xr = Xc[2];
        const float hx = xl + xm + xr, hxx = xl * xl + xm * xm + xr * xr, hxy = xl * yl + xm * ym + xr * yr;
        if (rr >= 2) {
          const float mu = (ax[f][ch] + hx) * inv9;
          st.mu_x[f][ch] = mu;
          st.sig_x[f][ch] = (axx[f][ch] + hxx) * inv9 - mu * mu;
          st.sig_xy[f][ch] = (axy[f][ch] + hxy) * inv9 - mu * st.mu_y[ch];
          st.xc[f][ch] = Xc[1 - PITCH2];
        }
        ax[f][ch] = bx[f][ch] + hx, bx[f][ch] = hx;
        axx[f][ch] = bxx[f][ch] + hxx, bxx[f][ch] = hxx;
        axy[f][ch] = bxy[f][ch] + hxy, bxy[f][ch] = hxy;
      }
    }
    if (rr >= 2) emit(rr - 2, st);
  }
}

__device__ __forceinline__ float ssim_raw(float mu_x, float sig_x, float sig_xy, float mu_y, float sig_y) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const float n = (2.f * mu_x * mu_y + C1) * (2.f * sig_xy + C2);
  const float dn = (mu_x * mu_x + mu_y * mu_y + C1) * (sig_x + sig_y + C2);
  return (1.f - n * rcp_nr(dn)) * 0.5f;
}

// weight with which full-resolution sample `dst` reads low-resolution sample `i` under bilinear
// up-sampling by 2^shift (align_corners=False): triangle kernel on the clamped source coordinate
__device__ __forceinline__ float tri_weight(int dst, int shift, int n_in, int i) {
  const float src = fminf(fmaxf(((float)dst + 0.5f) * (1.f / (float)(1 << shift)) - 0.5f, 0.f), (float)(n_in - 1));
  return fmaxf(0.f, 1.f - fabsf(src - (float)i));
}

constexpr int B_RUN = 3;                       // rows per stage-B thread
constexpr int B_THREADS = H1_W * (H1_H / B_RUN);   // 34 * 6 = 204

template <int MODE, int F, bool SAVED>
__global__ void __launch_bounds__(WP_THREADS, 2) warp_photo_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ CamConst cam;
  __shared__ float redT[WP_THREADS / 32][24];

  const dd_warp_desc& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * BT_H, c0 = blockIdx.x * BT_W;
  const int H = d.H, W = d.W;
  const size_t P = (size_t)H * W;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const bool automask = (d.flags & DD_FLAG_AUTOMASK) != 0;
  const float ssim_w = d.ssim_weight, l1_w = 1.f - d.ssim_weight;

  load_cam(&cam, d, b, tid);

  float* Y = smem + SM_Y;
  float* LID = smem + SM_LID;
  float* COEF = smem + SM_COEF;
  float* GT = smem + SM_GT;

  // ---- target (+ identity source) tiles over the 2-pixel halo -----------------------------------
  const float* tgt = d.target + (size_t)b * 3 * P;
  for (int i = tid; i < H2_W * H2_H; i += WP_THREADS) {
    const int hr = i / H2_W, hc = i - hr * H2_W;
    const int ri = r0 - 2 + hr, ci = c0 - 2 + hc;
    const bool used = ri >= -1 && ri <= H && ci >= -1 && ci <= W;
    const size_t o = (size_t)reflect1(max(min(ri, H), -1), H) * W + reflect1(max(min(ci, W), -1), W);
    const int so = hr * PITCH2 + hc;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) Y[ch * PLANE2 + so] = used ? __ldg(tgt + ch * P + o) : 0.f;
    if (automask) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float* src = d.source[f] + (size_t)b * 3 * P;
        float* X = smem + SM_X + f * 3 * PLANE2;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) X[ch * PLANE2 + so] = used ? __ldg(src + ch * P + o) : 0.f;
      }
    }
  }
  __syncthreads();

  if (automask) {   // identity candidate losses on the 1-pixel halo (level independent)
    if (tid < B_THREADS) {
      const int pc = tid % H1_W, pr0 = (tid / H1_W) * B_RUN;
      march_stats<F, B_RUN>(Y, smem + SM_X, pr0, pc, [&](int k, const PosStats& st) {
        float acc_s[2] = {0.f, 0.f}, acc_l[2] = {0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
          for (int f = 0; f < F; ++f) {
            acc_s[f] += fminf(fmaxf(ssim_raw(st.mu_x[f][ch], st.sig_x[f][ch], st.sig_xy[f][ch], st.mu_y[ch], st.sig_y[ch]), 0.f), 1.f);
            acc_l[f] += fabsf(st.yc[ch] - st.xc[f][ch]);
          }
#pragma unroll
        for (int f = 0; f < F; ++f)
          LID[f * CPLANE + (pr0 + k) * CPITCH + pc] = ssim_w * (acc_s[f] * (1.f / 3.f)) + l1_w * (acc_l[f] * (1.f / 3.f));
      });
    }
    __syncthreads();
  }

  float accT[2][12];
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < 12; ++k) accT[f][k] = 0.f;

  for (int si = 0; si < d.num_scales; ++si) {
    const int shift = d.scale[si];
    const int h = H >> shift, w = W >> shift;
    const size_t p_lo = (size_t)h * w;
    const float* disp = d.disp[si] + (size_t)b * p_lo;
    const float g_photo = __ldg(a.grad_sums + si * DD_NSUM + DD_SUM_PHOTO);
    const float g_coef = (-0.5f * (ssim_w / 3.f) * g_photo) / 9.f;   // d loss / d S per window tap

    // low-resolution patch of this level (consumed by stage C; the copies land under stages A and B)
    const float* patch = smem + SM_PATCH;
    const int ppw = (BT_W >> shift) + 2, pph = (BT_H >> shift) + 2;
    const int pbr = (r0 >> shift) - 1, pbc = (c0 >> shift) - 1;
    if (shift != 0) {
      for (int slot = tid; slot < ppw * pph; slot += WP_THREADS) {
        const int pr = slot / ppw, pc = slot - pr * ppw;
        const int off = min(max(pbr + pr, 0), h - 1) * w + min(max(pbc + pc, 0), w - 1);
        float* dst = smem + SM_PATCH + slot;
        cp_async4(dst, disp + off);
        if (MODE >= 1) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float* fl = d.flow[si][f] + (size_t)b * 3 * p_lo + off;
#pragma unroll
            for (int k = 0; k < 3; ++k) cp_async4(dst + (1 + 3 * f + k) * BP_PLANE, fl + k * p_lo);
            if (MODE == 2) cp_async4(dst + (7 + f) * BP_PLANE, d.mask[si][f] + (size_t)b * p_lo + off);
          }
        }
      }
    }

    // ---- stage A: warped frames over the 2-pixel halo --------------------------------------------
    if (SAVED) {
      for (int i = tid; i < H2_W * H2_H; i += WP_THREADS) {
        const int hr = i / H2_W, hc = i - hr * H2_W;
        const int ri = r0 - 2 + hr, ci = c0 - 2 + hc;
        const bool used = ri >= -1 && ri <= H && ci >= -1 && ci <= W;
        const size_t o = (size_t)reflect1(max(min(ri, H), -1), H) * W + reflect1(max(min(ci, W), -1), W);
        const int so = hr * PITCH2 + hc;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float* wsrc = a.warped_saved[si][f] + (size_t)b * 3 * P + o;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) smem[SM_X + (f * 3 + ch) * PLANE2 + so] = used ? __ldg(wsrc + ch * P) : 0.f;
        }
      }
      __syncthreads();
    } else {
      for (int i = tid; i < H2_W * H2_H; i += WP_THREADS) {
        const int hr = i / H2_W, hc = i - hr * H2_W;
        const int ri = r0 - 2 + hr, ci = c0 - 2 + hc;
        const bool used = ri >= -1 && ri <= H && ci >= -1 && ci <= W;
        const int so = hr * PITCH2 + hc;
        if (!used) {
  #pragma unroll
          for (int k = 0; k < 6; ++k) smem[SM_X + k * PLANE2 + so] = 0.f;
          continue;
        }
        const int r = reflect1(ri, H), c = reflect1(ci, W);
        const Taps ty = up_taps(r, shift, h), tx = up_taps(c, shift, w);
        PixelGeom pg;
        const float du = bilerp(disp, w, ty, tx);
        pg.depth = rcp_nr(a.min_disp + a.disp_range * du);
        const float u = (float)c, v = (float)r;
        pg.ray = {cam.iK[0] * u + cam.iK[1] * v + cam.iK[2], cam.iK[3] * u + cam.iK[4] * v + cam.iK[5],
                  cam.iK[6] * u + cam.iK[7] * v + cam.iK[8]};
        pg.Pc = {pg.depth * pg.ray.x, pg.depth * pg.ray.y, pg.depth * pg.ray.z};
  #pragma unroll
        for (int f = 0; f < F; ++f) {
          Vec3 cf = {0.f, 0.f, 0.f};
          float m = 1.f;
          if (MODE >= 1) {
            const float* fl = d.flow[si][f] + (size_t)b * 3 * p_lo;
            const float tsv = cam.ts[f];
            cf = {bilerp(fl, w, ty, tx) * tsv, bilerp(fl + p_lo, w, ty, tx) * tsv, bilerp(fl + 2 * p_lo, w, ty, tx) * tsv};
            if (MODE == 2) m = bilerp(d.mask[si][f] + (size_t)b * p_lo, w, ty, tx);
          }
          FrameGeom g;
          frame_geometry<MODE>(g, pg, &cam, f, cf, m, H, W, false);
          const Foot ft = footprint(unnormalise(g.gx, W), unnormalise(g.gy, H), H, W);
          const float* src = d.source[f] + (size_t)b * 3 * P;
  #pragma unroll
          for (int ch = 0; ch < 3; ++ch) smem[SM_X + (f * 3 + ch) * PLANE2 + so] = sample_plane(a.tex[f], src + ch * P, W, (b * 3 + ch) * H, ft);
        }
      }
      __syncthreads();
    }

    // ---- stage B: statistics, selection and coefficient maps on the 1-pixel halo -----------------
    if (tid < B_THREADS) {
      const int pc = tid % H1_W, pr0 = (tid / H1_W) * B_RUN;
      const int ci = c0 - 1 + pc;
      march_stats<F, B_RUN>(Y, smem + SM_X, pr0, pc, [&](int k, const PosStats& st) {
        const int pr = pr0 + k;
        const int ri = r0 - 1 + pr;
        float* co = COEF + pr * CPITCH + pc;
        if (ri < 0 || ri >= H || ci < 0 || ci >= W) {
          co[0] = __int_as_float(-1);
          return;
        }
        float vraw[2][3];
        float acc_s[2] = {0.f, 0.f}, acc_l[2] = {0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
          for (int f = 0; f < F; ++f) {
            vraw[f][ch] = ssim_raw(st.mu_x[f][ch], st.sig_x[f][ch], st.sig_xy[f][ch], st.mu_y[ch], st.sig_y[ch]);
            acc_s[f] += fminf(fmaxf(vraw[f][ch], 0.f), 1.f);
            acc_l[f] += fabsf(st.yc[ch] - st.xc[f][ch]);
          }
        // per-pixel argmin over {identity(+noise), warped} with first-index tie-break (Trainer.py:339-347)
        float best = 0.f;
        int arg = -1;
        bool first = true;
        const size_t o = (size_t)ri * W + ci;
        if (automask) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float v = LID[f * CPLANE + pr * CPITCH + pc];
            if (d.noise[si]) v += __ldg(d.noise[si] + ((size_t)b * F + f) * P + o) * 0.00001f;
            if (first || v < best) best = v, arg = -1, first = false;
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float v = ssim_w * (acc_s[f] * (1.f / 3.f)) + l1_w * (acc_l[f] * (1.f / 3.f));
          if (first || v < best) best = v, arg = f, first = false;
        }
        co[0] = __int_as_float(arg);
        if (arg >= 0) {
          const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
          const int fs = arg == 0 ? 0 : F - 1;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float mx = fs == 0 ? st.mu_x[0][ch] : st.mu_x[F - 1][ch];
            const float sx = fs == 0 ? st.sig_x[0][ch] : st.sig_x[F - 1][ch];
            const float sxy = fs == 0 ? st.sig_xy[0][ch] : st.sig_xy[F - 1][ch];
            const float vr = fs == 0 ? vraw[0][ch] : vraw[F - 1][ch];
            const bool gt = (vr >= 0.f) && (vr <= 1.f);   // torch.clamp passes the gradient on the closed interval
            const float my = st.mu_y[ch], sy = st.sig_y[ch];
            const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2;
            const float B1 = mx * mx + my * my + C1, B2 = sx + sy + C2;
            const float n = A1 * A2, dn = B1 * B2;
            const float inv_d = rcp_nr(dn);
            const float dS_dmu = (2.f * my * (A2 - A1) * dn - n * 2.f * mx * (B2 - B1)) * inv_d * inv_d;
            const float dS_dxx = -n * B1 * inv_d * inv_d;
            const float dS_dxy = 2.f * A1 * inv_d;
            const float G = gt ? g_coef : 0.f;
            co[(1 + ch * 3 + 0) * CPLANE] = G * dS_dmu;
            co[(1 + ch * 3 + 1) * CPLANE] = G * 2.f * dS_dxx;
            co[(1 + ch * 3 + 2) * CPLANE] = G * dS_dxy;
          }
        }
      });
    }
    cp_async_commit_wait();
    __syncthreads();

    // ---- stage C: per-pixel chain ----------------------------------------------------------------
    const bool want_disp = a.g.disp[si] != nullptr;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
      const int qr = warp * 2 + k, qc = lane;      // tile coordinates
      const int r = r0 + qr, c = c0 + qc;
      // box-sum of the coefficient maps over the 3x3 neighbourhood of windows (reflect fold-back weights);
      // grad x_f(q) = A_f + x_f(q) * B_f + y(q) * C_f per channel
      float cA[2][3], cB[2][3], cC[2][3];
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) cA[f][ch] = cB[f][ch] = cC[f][ch] = 0.f;
      bool any[2] = {false, false};
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int prr = r + dy;
        const float wr = ((r == 1 && prr == 0) || (r == H - 2 && prr == H - 1)) ? 2.f : 1.f;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int pcc = c + dx;
          const float* co = COEF + (qr + 1 + dy) * CPITCH + (qc + 1 + dx);
          const int sel = __float_as_int(co[0]);   // -1 outside the image or when an identity candidate won
          if (sel < 0) continue;
          const float wgt = wr * (((c == 1 && pcc == 0) || (c == W - 2 && pcc == W - 1)) ? 2.f : 1.f);
          const int fs = sel == 0 ? 0 : 1;
          any[fs] = true;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float va = wgt * co[(1 + ch * 3) * CPLANE], vb = wgt * co[(2 + ch * 3) * CPLANE], vc = wgt * co[(3 + ch * 3) * CPLANE];
            if (fs == 0) cA[0][ch] += va, cB[0][ch] += vb, cC[0][ch] += vc;
            else cA[1][ch] += va, cB[1][ch] += vb, cC[1][ch] += vc;
          }
        }
      }
      float gcol[2][3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float yq = Y[ch * PLANE2 + (qr + 2) * PITCH2 + qc + 2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float xq = smem[SM_X + (f * 3 + ch) * PLANE2 + (qr + 2) * PITCH2 + qc + 2];
          gcol[f][ch] = cA[f][ch] + xq * cB[f][ch] + yq * cC[f][ch];
        }
      }
      {   // L1 term of the centre pixel (Trainer.py:417-418)
        const int sel = __float_as_int(COEF[(qr + 1) * CPITCH + qc + 1]);
        if (sel >= 0) {
          const float gl = (l1_w / 3.f) * g_photo;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float yv = Y[ch * PLANE2 + (qr + 2) * PITCH2 + qc + 2];
            const float xv = smem[SM_X + (sel * 3 + ch) * PLANE2 + (qr + 2) * PITCH2 + qc + 2];
            const float dlt = yv - xv;
            const float sg = dlt > 0.f ? -gl : (dlt < 0.f ? gl : 0.f);   // d|y-x|/dx = -sign(y-x)
            if (sel == 0) gcol[0][ch] += sg; else gcol[1][ch] += sg;
          }
          any[sel == 0 ? 0 : 1] = true;
        }
      }

      // c_consistency upstream (scene-flow + mask phases): this pixel is one of the centre taps of
      // its 2^s block when it lies in the middle 2x2 (Trainer.py:284,385-386)
      bool centre = false;
      float cc_w = 0.f;
      int li = 0, lj = 0;
      if (MODE == 2) {
        if (shift == 0) {
          centre = true, cc_w = 1.f, li = r, lj = c;
        } else {
          const int half = 1 << (shift - 1), msk = (1 << shift) - 1;
          const int rr = r & msk, cr = c & msk;
          centre = (rr == half - 1 || rr == half) && (cr == half - 1 || cr == half);
          cc_w = 0.25f, li = r >> shift, lj = c >> shift;
        }
      }

      float gd_up = 0.f;                      // d/d disp_up (summed over frames)
      float gcf_up[2][3], gm_up[2];
#pragma unroll
      for (int f = 0; f < 2; ++f) gcf_up[f][0] = gcf_up[f][1] = gcf_up[f][2] = gm_up[f] = 0.f;

      const Taps ty = up_taps(r, shift, h), tx = up_taps(c, shift, w);
      const int py0 = ty.i0 - pbr, py1 = ty.i1 - pbr, px0 = tx.i0 - pbc, px1 = tx.i1 - pbc;   // patch-relative taps
      PixelGeom pg;
      {
        const float du = shift == 0 ? __ldg(disp + (size_t)r * W + c) : patch_bilerp(patch, ppw, py0, py1, px0, px1, ty.l, tx.l);
        pg.depth = rcp_nr(a.min_disp + a.disp_range * du);
        const float u = (float)c, v = (float)r;
        pg.ray = {cam.iK[0] * u + cam.iK[1] * v + cam.iK[2], cam.iK[3] * u + cam.iK[4] * v + cam.iK[5],
                  cam.iK[6] * u + cam.iK[7] * v + cam.iK[8]};
        pg.Pc = {pg.depth * pg.ray.x, pg.depth * pg.ray.y, pg.depth * pg.ray.z};
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float g_cc = 0.f;
        if (MODE == 2 && centre) g_cc = __ldg(a.grad_sums + si * DD_NSUM + DD_SUM_CONSIST0 + f);
        const bool cc_live = (MODE == 2) && centre && (g_cc != 0.f);
        if (!any[f] && !cc_live) continue;
        Vec3 cf = {0.f, 0.f, 0.f};
        float m = 1.f;
        if (MODE >= 1) {
          const float tsv = cam.ts[f];
          if (shift == 0) {
            const float* fl = d.flow[si][f] + (size_t)b * 3 * p_lo + (size_t)r * W + c;
            cf = {__ldg(fl) * tsv, __ldg(fl + p_lo) * tsv, __ldg(fl + 2 * p_lo) * tsv};
            if (MODE == 2) m = __ldg(d.mask[si][f] + (size_t)b * p_lo + (size_t)r * W + c);
          } else {
            const float* fp = patch + (1 + 3 * f) * BP_PLANE;
            cf = {patch_bilerp(fp, ppw, py0, py1, px0, px1, ty.l, tx.l) * tsv,
                  patch_bilerp(fp + BP_PLANE, ppw, py0, py1, px0, px1, ty.l, tx.l) * tsv,
                  patch_bilerp(fp + 2 * BP_PLANE, ppw, py0, py1, px0, px1, ty.l, tx.l) * tsv};
            if (MODE == 2) m = patch_bilerp(patch + (7 + f) * BP_PLANE, ppw, py0, py1, px0, px1, ty.l, tx.l);
          }
        }
        FrameGeom g;
        frame_geometry<MODE>(g, pg, &cam, f, cf, m, H, W, false);
        Vec3 gPin = {0.f, 0.f, 0.f};
        if (any[f]) {
          const float uix = unnormalise(g.gx, W), uiy = unnormalise(g.gy, H);
          const Foot ft = footprint(uix, uiy, H, W);
          const float* src = d.source[f] + (size_t)b * 3 * P;
          float gix = 0.f, giy = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const Quad q = gather4(a.tex[f], src + ch * P, W, (b * 3 + ch) * H, ft);
            gix += gcol[f][ch] * ((q.ne - q.nw) * ft.wy0 + (q.se - q.sw) * ft.wy1);
            giy += gcol[f][ch] * ((q.sw - q.nw) * ft.wx0 + (q.se - q.ne) * ft.wx1);
          }
          const float gpx = ft.live_x ? gix : 0.f, gpy = ft.live_y ? giy : 0.f;   // GridSampler.cuh:64-80
          const float iz = g.pr.iz;
          const float gc0 = gpx * iz, gc1 = gpy * iz, gc2 = -(gpx * g.pr.px + gpy * g.pr.py) * iz;
          const float* K = cam.K;
          const Vec3 gX = {K[0] * gc0 + K[4] * gc1 + K[8] * gc2, K[1] * gc0 + K[5] * gc1 + K[9] * gc2,
                           K[2] * gc0 + K[6] * gc1 + K[10] * gc2};
          if (MODE != 1) {
            const float* T = cam.T[f];
            accT[f][0] += gX.x * g.Pin.x, accT[f][1] += gX.x * g.Pin.y, accT[f][2] += gX.x * g.Pin.z, accT[f][3] += gX.x;
            accT[f][4] += gX.y * g.Pin.x, accT[f][5] += gX.y * g.Pin.y, accT[f][6] += gX.y * g.Pin.z, accT[f][7] += gX.y;
            accT[f][8] += gX.z * g.Pin.x, accT[f][9] += gX.z * g.Pin.y, accT[f][10] += gX.z * g.Pin.z, accT[f][11] += gX.z;
            gPin = {T[0] * gX.x + T[4] * gX.y + T[8] * gX.z, T[1] * gX.x + T[5] * gX.y + T[9] * gX.z,
                    T[2] * gX.x + T[6] * gX.y + T[10] * gX.z};
          } else {
            gPin = gX;
          }
        }
        Vec3 gPc = gPin;
        if (MODE == 1) {
          gcf_up[f][0] = gPin.x, gcf_up[f][1] = gPin.y, gcf_up[f][2] = gPin.z;
        }
        if (MODE == 2) {
          gm_up[f] = gPin.x * g.res.x + gPin.y * g.res.y + gPin.z * g.res.z;
          Vec3 gres = {gPin.x * m, gPin.y * m, gPin.z * m};
          if (cc_live) {
            const size_t ol = (size_t)li * w + lj;
            const float valid = __ldg(disp + ol) > d.mask_disp_thrd ? 1.f : 0.f;
            const float ms = __ldg(d.mask[si][f] + (size_t)b * p_lo + ol);
            const float kcc = g_cc * valid * (1.f - ms) * cc_w;
            float rs[3] = {g.res.x, g.res.y, g.res.z};
            if (shift != 0) {
              const float* rsv = a.resid_saved[si][f] + (size_t)b * 3 * p_lo + ol;
              rs[0] = __ldg(rsv), rs[1] = __ldg(rsv + p_lo), rs[2] = __ldg(rsv + 2 * p_lo);
            }
            gres.x += kcc * (rs[0] > 0.f ? 1.f : (rs[0] < 0.f ? -1.f : 0.f));
            gres.y += kcc * (rs[1] > 0.f ? 1.f : (rs[1] < 0.f ? -1.f : 0.f));
            gres.z += kcc * (rs[2] > 0.f ? 1.f : (rs[2] < 0.f ? -1.f : 0.f));
          }
          gcf_up[f][0] = gres.x, gcf_up[f][1] = gres.y, gcf_up[f][2] = gres.z;
          // ego = (T @ (Pc,1))[:3] - Pc ; res = cf - ego
          const Vec3 ge = {-gres.x, -gres.y, -gres.z};
          const float* T = cam.T[f];
          accT[f][0] += ge.x * pg.Pc.x, accT[f][1] += ge.x * pg.Pc.y, accT[f][2] += ge.x * pg.Pc.z, accT[f][3] += ge.x;
          accT[f][4] += ge.y * pg.Pc.x, accT[f][5] += ge.y * pg.Pc.y, accT[f][6] += ge.y * pg.Pc.z, accT[f][7] += ge.y;
          accT[f][8] += ge.z * pg.Pc.x, accT[f][9] += ge.z * pg.Pc.y, accT[f][10] += ge.z * pg.Pc.z, accT[f][11] += ge.z;
          gPc.x += T[0] * ge.x + T[4] * ge.y + T[8] * ge.z - ge.x;
          gPc.y += T[1] * ge.x + T[5] * ge.y + T[9] * ge.z - ge.y;
          gPc.z += T[2] * ge.x + T[6] * ge.y + T[10] * ge.z - ge.z;
        }
        const float g_depth = pg.ray.x * gPc.x + pg.ray.y * gPc.y + pg.ray.z * gPc.z;
        gd_up += -a.disp_range * pg.depth * pg.depth * g_depth;
        if (MODE >= 1) {
          const float tsv = cam.ts[f];
          gcf_up[f][0] *= tsv, gcf_up[f][1] *= tsv, gcf_up[f][2] *= tsv;
        }
      }

      // hand the full-resolution gradients to stage D (level 0: identity up-sampling, store directly)
      const int q = qr * BT_W + qc;
      if (shift == 0) {
        const size_t o = (size_t)r * W + c;
        if (want_disp) a.g.disp[si][(size_t)b * P + o] = gd_up;
        if (MODE >= 1) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            if (a.g.flow[si][f]) {
              float* go = a.g.flow[si][f] + (size_t)b * 3 * P + o;
              go[0] = gcf_up[f][0], go[P] = gcf_up[f][1], go[2 * P] = gcf_up[f][2];
            }
            if (MODE == 2 && a.g.mask[si][f]) a.g.mask[si][f][(size_t)b * P + o] = gm_up[f];
          }
        }
      } else {
        GT[q] = gd_up;
        if (MODE >= 1) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            GT[(1 + f * 3 + 0) * GPLANE + q] = gcf_up[f][0];
            GT[(1 + f * 3 + 1) * GPLANE + q] = gcf_up[f][1];
            GT[(1 + f * 3 + 2) * GPLANE + q] = gcf_up[f][2];
            if (MODE == 2) GT[(7 + f) * GPLANE + q] = gm_up[f];
          }
        }
      }
    }
    __syncthreads();

    // ---- stage D: transposed bilinear up-sampling (levels > 0), rows then columns ---------------------
    if (shift != 0) {
      const int tl_h = (BT_H >> shift) + 2, tl_w = (BT_W >> shift) + 2;
      const int narr = MODE == 0 ? 1 : (MODE == 1 ? 7 : 9);
      const int win = 2 << shift, half = 1 << (shift - 1);
      float* V = COEF;   // [narr][tl_h][32] -- the coefficient maps are dead after stage C
      // pass 1: V[a][li][c] = sum over the tile rows that read low-res row gi of wy * GT[a][row][c];
      // thread = (li, c): the row weights are shared by all `narr` gradient planes
      for (int t = tid; t < tl_h * BT_W; t += WP_THREADS) {
        const int cx = t & (BT_W - 1), li = t >> 5;
        const int gi = (r0 >> shift) - 1 + li;
        float acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.f;
        if (gi >= 0 && gi < h) {
          const int rs = (gi << shift) - half;
          const int e0 = max(0, r0 - rs), e1 = min(win, r0 + BT_H - rs);
          for (int e = e0; e < e1; ++e) {
            const int rr = rs + e;
            const float wy = tri_weight(rr, shift, h, gi);
            const float* gp = GT + (rr - r0) * BT_W + cx;
#pragma unroll
            for (int k = 0; k < 9; ++k)
              if (k < narr) acc[k] += wy * gp[k * GPLANE];
          }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (k < narr) V[(k * tl_h + li) * BT_W + cx] = acc[k];
      }
      __syncthreads();
      // pass 2: out[a][li][lj] = sum over the tile columns that read low-res column gj of wx * V[a][li][col];
      // thread = (li, lj)
      for (int t = tid; t < tl_h * tl_w; t += WP_THREADS) {
        const int li = t / tl_w, lj = t - li * tl_w;
        const int gi = (r0 >> shift) - 1 + li, gj = (c0 >> shift) - 1 + lj;
        if (gi < 0 || gi >= h || gj < 0 || gj >= w) continue;
        const int cs = (gj << shift) - half;
        const int e0 = max(0, c0 - cs), e1 = min(win, c0 + BT_W - cs);
        float acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.f;
        for (int e = e0; e < e1; ++e) {
          const int cc = cs + e;
          const float wx = tri_weight(cc, shift, w, gj);
          const float* vp = V + li * BT_W + (cc - c0);
#pragma unroll
          for (int k = 0; k < 9; ++k)
            if (k < narr) acc[k] += wx * vp[k * tl_h * BT_W];
        }
        const size_t ol = (size_t)gi * w + gj;
        if (want_disp && acc[0] != 0.f) atomicAdd(a.g.disp[si] + (size_t)b * p_lo + ol, acc[0]);
        if (MODE >= 1) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            if (a.g.flow[si][f]) {
              float* gf = a.g.flow[si][f] + (size_t)b * 3 * p_lo + ol;
#pragma unroll
              for (int ch = 0; ch < 3; ++ch)
                if (acc[1 + f * 3 + ch] != 0.f) atomicAdd(gf + ch * p_lo, acc[1 + f * 3 + ch]);
            }
            if (MODE == 2 && a.g.mask[si][f] && acc[7 + f] != 0.f) atomicAdd(a.g.mask[si][f] + (size_t)b * p_lo + ol, acc[7 + f]);
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- pose gradient: per-CTA partial [f][12] ------------------------------------------------------
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(accT[f][k]);
      if (lane == 0) redT[warp][f * 12 + k] = v;
    }
  __syncthreads();
  if (tid < 24) {
    float v = 0.f;
#pragma unroll
    for (int wi = 0; wi < WP_THREADS / 32; ++wi) v += redT[wi][tid];
    a.partial_T[(size_t)cta * 24 + tid] = v;
  }
}

// grad_T[f][b][i][j] = sum over the image's CTAs (deterministic order); row 3 is zero because
// K[:3,3] == 0 for pinhole intrinsics (datasets/base_dataset.py:154-163).
__global__ void finalize_T_kernel(const float* __restrict__ partial, float* __restrict__ gT0, float* __restrict__ gT1,
                                  int ctas_per_image) {
  const int b = blockIdx.x;
  const int k = threadIdx.x;   // 0..31 : f*16 + entry
  const int f = k >> 4, e = k & 15;
  float* out = f == 0 ? gT0 : gT1;
  if (!out) return;
  double acc = 0.0;
  if (e < 12)
    for (int i = 0; i < ctas_per_image; ++i) acc += (double)partial[((size_t)b * ctas_per_image + i) * 24 + f * 12 + e];
  out[b * 16 + e] = (float)acc;
}

int validate_desc(const dd_warp_desc* d);

template <int MODE, int F, bool SAVED>
static int launch_bwd(const BwdArgs& args, dim3 grid, cudaStream_t st) {
  auto kern = warp_photo_bwd_kernel<MODE, F, SAVED>;
  const size_t smem_bytes = SM_TOTAL * sizeof(float);
  DD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  kern<<<grid, WP_THREADS, smem_bytes, st>>>(args); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int warp_photo_bwd_impl(const dd_warp_desc* desc, const float* grad_sums, const dd_warp_grads* grads,
                        const dd_warp_aux* saved, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  int rc = validate_desc(desc);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(grad_sums != nullptr && grads != nullptr, "dd_warp_photo_bwd: grad_sums / grads is NULL");
  const size_t need = dd_warp_photo_workspace_bytes(desc);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("dd_warp_photo_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DD_ERR_WORKSPACE;
  }
  const int mode = (desc->flags & DD_FLAG_CMPFLOW) ? ((desc->flags & DD_FLAG_MOTMASK) ? 2 : 1) : 0;
  BwdArgs args;
  memset(&args, 0, sizeof(args));
  args.d = *desc;
  args.g = *grads;
  args.grad_sums = grad_sums;
  args.partial_T = reinterpret_cast<float*>(workspace);
  args.min_disp = 1.f / desc->max_depth;
  args.disp_range = 1.f / desc->min_depth - 1.f / desc->max_depth;
  for (int f = 0; f < DD_MAX_FRAMES; ++f) args.tex[f] = f < desc->num_frames ? source_texture(desc->source[f], desc->B, desc->H, desc->W) : 0;
  for (int s = 0; s < desc->num_scales; ++s) {
    const size_t p_lo = (size_t)(desc->H >> desc->scale[s]) * (desc->W >> desc->scale[s]);
    for (int f = 0; f < desc->num_frames; ++f) {
      if (mode == 2 && desc->scale[s] != 0) {
        DD_REQUIRE(saved && saved->resid[s][f], "dd_warp_photo_bwd: saved->resid[%d][%d] (forward by-product) is required", s, f);
        args.resid_saved[s][f] = saved->resid[s][f];
      }
      if (mode == 2 && grads->mask[s][0] && grads->mask[s][1])
        DD_REQUIRE(grads->mask[s][0] != grads->mask[s][1], "grads.mask[%d][0] and [1] must not alias", s);
    }
    // levels > 0 accumulate with atomics -> zero first; level 0 is fully overwritten by plain stores
    if (desc->scale[s] != 0) {
      if (grads->disp[s]) DD_CHECK_CUDA(cudaMemsetAsync(grads->disp[s], 0, desc->B * p_lo * sizeof(float), st));
      for (int f = 0; f < desc->num_frames; ++f) {
        if (mode >= 1 && grads->flow[s][f]) DD_CHECK_CUDA(cudaMemsetAsync(grads->flow[s][f], 0, desc->B * 3 * p_lo * sizeof(float), st));
        if (mode == 2 && grads->mask[s][f]) DD_CHECK_CUDA(cudaMemsetAsync(grads->mask[s][f], 0, desc->B * p_lo * sizeof(float), st));
      }
    }
  }
  const dim3 grid(desc->W / BT_W, desc->H / BT_H, desc->B);
  const int F = desc->num_frames;
  bool have_warped = saved != nullptr;
  for (int s = 0; s < desc->num_scales && have_warped; ++s)
    for (int f = 0; f < F; ++f) {
      if (!saved->warped[s][f]) have_warped = false;
      else args.warped_saved[s][f] = saved->warped[s][f];
    }
#define DD_BWD(M, FF)                                                      \
  rc = have_warped ? launch_bwd<M, FF, true>(args, grid, st) : launch_bwd<M, FF, false>(args, grid, st)
  if (mode == 0 && F == 2) DD_BWD(0, 2);
  else if (mode == 1 && F == 2) DD_BWD(1, 2);
  else if (mode == 2 && F == 2) DD_BWD(2, 2);
  else if (mode == 0 && F == 1) DD_BWD(0, 1);
  else if (mode == 1 && F == 1) DD_BWD(1, 1);
  else DD_BWD(2, 1);
#undef DD_BWD
  if (rc != DD_OK) return rc;
  if (grads->T[0] || grads->T[1]) {
    finalize_T_kernel<<<desc->B, 32, 0, st>>>(args.partial_T, grads->T[0], F > 1 ? grads->T[1] : nullptr,
                                              (int)(grid.x * grid.y)); dd::count_launches(1);
    DD_CHECK_CUDA(cudaGetLastError());
  }
  return DD_OK;
}

}  // namespace dd
