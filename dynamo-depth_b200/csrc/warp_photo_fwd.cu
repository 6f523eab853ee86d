// Fused view synthesis + photometric loss, forward (dd_warp_photo_fwd).
//
// One CTA owns a 32x32 tile of one image and walks every pyramid level and source frame:
//   stage A  per halo pixel (34x34, ReflectionPad2d(1) folded into the index map): bilinear
//            up-sample of disp_s (and flow_s / mask_s), disp->depth, back-projection, rigid or
//            scene-flow transform, projection, border-clamped bilinear gather of the source frame
//            -> warped colour tile in shared memory (target tile staged once per CTA);
//   stage B  3x3 SSIM statistics by a register ring marching down each column (horizontal taps
//            from shared memory), L1, 0.85/0.15 mix, per-pixel min over {identity, warped}
//            candidates, block reduction of the selected loss;
//   stage C  (scene-flow phases) 2^s-block centre averages of the residual flow and of
//            sample_ego - sample_complete -> c_consistency sum and disp_mag map.
// HBM traffic per image and level: target + 2 sources once (colours of other levels hit L2),
// disp_s once; nothing full-resolution is written unless an aux output is requested.
#include "warp_photo.cuh"

namespace dd {

constexpr int HALO1 = TILE + 2;   // 34
constexpr int PITCH1 = 35;   // odd pitch: conflict-free rows; 3 CTAs x (dynamic + static + 1 KB) must stay <= 196 KB (carve-out step)
constexpr int PLANE1 = HALO1 * PITCH1;

struct FwdArgs {
  dd_warp_desc d;
  dd_warp_aux aux;
  float* partial;   // [num_scales*DD_NSUM][num_ctas]
  float min_disp, disp_range;
  int has_aux;
  cudaTextureObject_t tex[DD_MAX_FRAMES];   // source frames as textures (0: gather with plain loads)
};

// shared memory carve-up (floats): Y[3][34][35] | XI[3][34][35][2] | lo[2][5][16][16] (flow modes: low-resolution
// accumulators of the 2^s-block centre averages, filled with shared-memory atomics)
__device__ __forceinline__ float* smem_Y(float* s) { return s; }
// warped (or identity) source tiles, frame-interleaved: element (ch, pos, f) at XI[(ch*PLANE1 + pos)*2 + f], so that the
// two frames of a pixel are one 64-bit word = one FFMA2 operand of the SSIM statistics
__device__ __forceinline__ float* smem_XI(float* s) { return s + 3 * PLANE1; }
__device__ __forceinline__ float* smem_lo(float* s) { return s + 9 * PLANE1; }
constexpr int LO_PLANE = (TILE / 2) * (TILE / 2);   // 256 low-res pixels per tile at level 1 (the largest staged level)
// levels > 0: the low-resolution disp / flow / mask texels every halo pixel of the tile interpolates from, staged once
// per level with cp.async ((TILE >> s) + 2)^2 texels per plane; 9 planes: disp | flow f0 xyz | flow f1 xyz | mask f0 | mask f1
constexpr int PATCH_MAXW = TILE / 2 + 2;              // 18 (level 1)
constexpr int PATCH_PLANE = PATCH_MAXW * PATCH_MAXW;  // 324
__device__ __forceinline__ float* smem_patch(float* s, int mode) { return s + 9 * PLANE1 + (mode >= 1 ? 2 * 5 * LO_PLANE : 0); }


struct SsimOut {
  float L[2][4];   // per frame, per row of the thread's 4-row run
};

// Stage B: thread = (column lane, 4-row run).  Returns the mixed reprojection loss
// ssim_w*mean_c(SSIM) + l1_w*mean_c|y-x| (Trainer.py:413-423, tools.py:243-257) per frame.  The statistics of the two
// source frames are carried as packed pairs (frame 0, frame 1): sums, products and the SSIM rational run as
// FADD2 / FMUL2 / FFMA2, i.e. one issue slot for both frames; target-only statistics stay scalar.
template <int F>
__device__ __forceinline__ void ssim_l1_run(const float* __restrict__ Y, const float* __restrict__ XI, int lane, int row0,
                                            float ssim_w, float l1_w, SsimOut& out) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const float inv9 = 1.f / 9.f, inv3 = 1.f / 3.f;   // window / channel means as multiplies (<= 1 ulp from the divisions)
  const f32x2 inv9_2 = pack2(inv9, inv9), ninv9_2 = pack2(-inv9, -inv9);
  const f32x2 C1_2 = pack2(C1, C1), C2_2 = pack2(C2, C2), two_2 = pack2(2.f, 2.f), one_2 = pack2(1.f, 1.f), half_2 = pack2(0.5f, 0.5f);
  float ssim_acc[2][4], l1_acc[2][4];
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k) ssim_acc[f][k] = 0.f, l1_acc[f][k] = 0.f;

#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* Yc = Y + ch * PLANE1;
    const float* Xc = XI + ch * PLANE1 * 2;
    float ay = 0.f, by = 0.f, ayy = 0.f, byy = 0.f;
    f32x2 ax = 0ull, bx = 0ull, axx = 0ull, bxx = 0ull, axy = 0ull, bxy = 0ull;
    float yc_prev = 0.f;
    f32x2 xc_prev = 0ull;
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
      const int o = (row0 + rr) * PITCH1 + lane;
      const float yl = Yc[o], yc = Yc[o + 1], yr = Yc[o + 2];
      const f32x2 xl = *reinterpret_cast<const f32x2*>(Xc + 2 * o), xc = *reinterpret_cast<const f32x2*>(Xc + 2 * o + 2),
                  xr = *reinterpret_cast<const f32x2*>(Xc + 2 * o + 4);
      const float hy = yl + yc + yr;
      const float hyy = yl * yl + yc * yc + yr * yr;
      const f32x2 hx = add2(add2(xl, xc), xr);
      const f32x2 hxx = fma2(xr, xr, fma2(xc, xc, mul2(xl, xl)));
      const f32x2 hxy = fma2(xr, pack2(yr, yr), fma2(xc, pack2(yc, yc), mul2(xl, pack2(yl, yl))));
      if (rr >= 2) {
        const int k = rr - 2;   // output row row0+k, centre halo row row0+k+1 == previous iteration
        const float mu_y = (ay + hy) * inv9;
        const float e_yy = (ayy + hyy) * inv9;
        const float sig_y = e_yy - mu_y * mu_y;
        const f32x2 sx = add2(ax, hx);
        const f32x2 mu_x = mul2(sx, inv9_2), nmu_x = mul2(sx, ninv9_2);
        const f32x2 nsig_x = fma2(mu_x, mu_x, mul2(add2(axx, hxx), ninv9_2));          // -(E[xx] - mu_x^2)
        const f32x2 sig_xy = fma2(nmu_x, pack2(mu_y, mu_y), mul2(add2(axy, hxy), inv9_2));
        const f32x2 n = mul2(fma2(mu_x, pack2(2.f * mu_y, 2.f * mu_y), C1_2), fma2(sig_xy, two_2, C2_2));
        const float dy1 = mu_y * mu_y + C1, ndy2 = -(sig_y + C2);
        const f32x2 ndn = mul2(fma2(mu_x, mu_x, pack2(dy1, dy1)), add2(nsig_x, pack2(ndy2, ndy2)));   // -dn
        float nd0, nd1;
        unpack2(ndn, nd0, nd1);
        const f32x2 nr = pack2(rcp_nr(nd0), rcp_nr(nd1));                                 // -1/dn
        float s0, s1;
        unpack2(mul2(fma2(n, nr, one_2), half_2), s0, s1);                                // (1 - n/dn)/2
        ssim_acc[0][k] += __saturatef(s0);
        ssim_acc[1][k] += __saturatef(s1);
        float d0, d1;
        unpack2(add2(xc_prev, pack2(-yc_prev, -yc_prev)), d0, d1);
        l1_acc[0][k] += fabsf(d0);
        l1_acc[1][k] += fabsf(d1);
      }
      ay = by + hy, by = hy, ayy = byy + hyy, byy = hyy;
      yc_prev = yc;
      ax = add2(bx, hx), bx = hx;
      axx = add2(bxx, hxx), bxx = hxx;
      axy = add2(bxy, hxy), bxy = hxy;
      xc_prev = xc;
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k) out.L[f][k] = ssim_w * (ssim_acc[f][k] * inv3) + l1_w * (l1_acc[f][k] * inv3);
}

template <int MODE, int F>
__global__ void __launch_bounds__(WP_THREADS, (MODE == 0 ? 4 : 3)) warp_photo_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ CamConst cam;
  __shared__ float red[WP_THREADS / 32][8];

  const dd_warp_desc& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * TILE, c0 = blockIdx.x * TILE;
  const int H = d.H, W = d.W;
  const size_t P = (size_t)H * W;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int num_ctas = gridDim.x * gridDim.y * gridDim.z;
  const bool automask = (d.flags & DD_FLAG_AUTOMASK) != 0;

  load_cam(&cam, d, b, tid);

  float* Y = smem_Y(smem);
  // ---- target tile (+ identity source tiles) ------------------------------------------------
  const float* tgt = d.target + (size_t)b * 3 * P;
  for (int i = tid; i < HALO1 * HALO1; i += WP_THREADS) {
    const int hr = i / HALO1, hc = i - hr * HALO1;
    const int r = reflect1(r0 - 1 + hr, H), c = reflect1(c0 - 1 + hc, W);
    const size_t o = (size_t)r * W + c;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) Y[ch * PLANE1 + hr * PITCH1 + hc] = __ldg(tgt + ch * P + o);
    if (automask) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float* src = d.source[f] + (size_t)b * 3 * P;
        float* X = smem_XI(smem) + f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) X[(ch * PLANE1 + hr * PITCH1 + hc) * 2] = __ldg(src + ch * P + o);
      }
    }
  }
  __syncthreads();

  // issues the cp.async copies of level si's low-resolution patch (no-op for full-resolution levels)
  auto stage_patch = [&](int si) {
    const int shift = d.scale[si];
    if (shift == 0) return;
    const int h = H >> shift, w = W >> shift, pw = (TILE >> shift) + 2;
    const int br = (r0 >> shift) - 1, bc = (c0 >> shift) - 1;
    const size_t p_lo = (size_t)h * w;
    float* patch = smem_patch(smem, MODE);
    // thread = patch texel(s); every plane of that texel shares the (clamped) source offset
    for (int slot = tid; slot < pw * pw; slot += WP_THREADS) {
      const int pr = slot / pw, pc = slot - pr * pw;
      const int off = min(max(br + pr, 0), h - 1) * w + min(max(bc + pc, 0), w - 1);
      float* dst = patch + slot;
      cp_async4(dst, d.disp[si] + (size_t)b * p_lo + off);
      if (MODE >= 1) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float* fl = d.flow[si][f] + (size_t)b * 3 * p_lo + off;
#pragma unroll
          for (int k = 0; k < 3; ++k) cp_async4(dst + (1 + 3 * f + k) * PATCH_PLANE, fl + k * p_lo);
          if (MODE == 2) cp_async4(dst + (1 + 3 * F + f) * PATCH_PLANE, d.mask[si][f] + (size_t)b * p_lo + off);
        }
      }
    }
  };
  stage_patch(0);

  const float l1_w = 1.f - d.ssim_weight;
  SsimOut ident;
  if (automask) {   // identity reprojection losses (Trainer.py:327-333), level independent
    ssim_l1_run<F>(Y, smem_XI(smem), lane, warp * 4, d.ssim_weight, l1_w, ident);
    __syncthreads();
  }

  for (int si = 0; si < d.num_scales; ++si) {
    const int shift = d.scale[si];
    const int h = H >> shift, w = W >> shift;
    const size_t p_lo = (size_t)h * w;
    const float* disp = d.disp[si] + (size_t)b * p_lo;

    const float* patch = smem_patch(smem, MODE);
    const int pw = (TILE >> shift) + 2, pbr = (r0 >> shift) - 1, pbc = (c0 >> shift) - 1;
    cp_async_commit_wait();
    __syncthreads();   // this level's low-resolution patch is in shared memory

    float s_cc[2] = {0.f, 0.f}, s_mag[2] = {0.f, 0.f};
    // ---- stage A: warp every halo pixel of both frames ---------------------------------------
    // Interior pixels: thread (warp w, lane l) owns column l, rows 4w .. 4w+3 (the mapping of stage B): the 2x2 centre taps of a
    // 2^s block that bilinear down-sampling averages are then either in this thread (rows) or one shuffle away (columns), so the
    // low-resolution by-products are reduced in registers -- no shared-memory atomics (CAS loops on this architecture).
    // Pass 4 covers the 132 pixels of the halo ring with the first 132 threads.
#pragma unroll 1
    for (int pass = 0; pass < 5; ++pass) {
      int hr, hc;
      if (pass < 4) {
        hr = 1 + 4 * warp + pass, hc = 1 + lane;
      } else {
        if (tid >= 2 * HALO1 + 2 * TILE) break;
        if (tid < HALO1) hr = 0, hc = tid;
        else if (tid < 2 * HALO1) hr = HALO1 - 1, hc = tid - HALO1;
        else if (tid < 2 * HALO1 + TILE) hr = 1 + (tid - 2 * HALO1), hc = 0;
        else hr = 1 + (tid - 2 * HALO1 - TILE), hc = HALO1 - 1;
      }
      const bool interior = pass < 4;
      const int r = reflect1(r0 - 1 + hr, H), c = reflect1(c0 - 1 + hc, W);
      const Taps ty = up_taps(r, shift, h), tx = up_taps(c, shift, w);
      // patch-relative tap indices (levels > 0)
      const int py0 = ty.i0 - pbr, py1 = ty.i1 - pbr, px0 = tx.i0 - pbc, px1 = tx.i1 - pbc;
      PixelGeom pg;
      {
        const float du = shift == 0 ? __ldg(disp + (size_t)r * W + c)
                                    : patch_bilerp(patch, pw, py0, py1, px0, px1, ty.l, tx.l);   // Trainer.py:225
        pg.depth = rcp_nr(a.min_disp + a.disp_range * du);            // tools.py:291-298
        const float u = (float)c, v = (float)r;
        pg.ray = {cam.iK[0] * u + cam.iK[1] * v + cam.iK[2], cam.iK[3] * u + cam.iK[4] * v + cam.iK[5],
                  cam.iK[6] * u + cam.iK[7] * v + cam.iK[8]};            // tools.py:193
        pg.Pc = {pg.depth * pg.ray.x, pg.depth * pg.ray.y, pg.depth * pg.ray.z};   // tools.py:194
      }
      const size_t o = (size_t)r * W + c;
      if (interior && a.has_aux && a.aux.depth[si]) a.aux.depth[si][(size_t)b * P + o] = pg.depth;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        Vec3 cf = {0.f, 0.f, 0.f};
        float m = 1.f;
        if (MODE >= 1) {
          const float tsv = cam.ts[f];
          if (shift == 0) {
            const float* fl = d.flow[si][f] + (size_t)b * 3 * p_lo + (size_t)r * W + c;
            cf = {__ldg(fl) * tsv, __ldg(fl + p_lo) * tsv, __ldg(fl + 2 * p_lo) * tsv};               // Trainer.py:251
            if (MODE == 2) m = __ldg(d.mask[si][f] + (size_t)b * p_lo + (size_t)r * W + c);            // Trainer.py:242
          } else {
            const float* fp = patch + (1 + 3 * f) * PATCH_PLANE;
            cf = {patch_bilerp(fp, pw, py0, py1, px0, px1, ty.l, tx.l) * tsv,
                  patch_bilerp(fp + PATCH_PLANE, pw, py0, py1, px0, px1, ty.l, tx.l) * tsv,
                  patch_bilerp(fp + 2 * PATCH_PLANE, pw, py0, py1, px0, px1, ty.l, tx.l) * tsv};
            if (MODE == 2) m = patch_bilerp(patch + (1 + 3 * F + f) * PATCH_PLANE, pw, py0, py1, px0, px1, ty.l, tx.l);
          }
        }
        FrameGeom g;
        frame_geometry<MODE>(g, pg, &cam, f, cf, m, H, W, interior);
        const Foot ft = footprint(unnormalise(g.gx, W), unnormalise(g.gy, H), H, W);
        const float* src = d.source[f] + (size_t)b * 3 * P;
        float* X = smem_XI(smem) + f;
        float col[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          col[ch] = sample_plane(a.tex[f], src + ch * P, W, (b * 3 + ch) * H, ft);   // Trainer.py:281
          X[(ch * PLANE1 + hr * PITCH1 + hc) * 2] = col[ch];
        }
        if (interior) {
          if (MODE >= 1) {
            if (shift == 0) {   // level 0: the "down-sampled" by-products are the pixel's own values
              const float mag = g.dsx * g.dsx + g.dsy * g.dsy;                 // Trainer.py:396
              s_mag[f] += mag;
              if (a.has_aux && a.aux.mag[si][f]) a.aux.mag[si][f][(size_t)b * P + o] = mag;
              if (a.has_aux && a.aux.resid[si][f]) {
                float* ro = a.aux.resid[si][f] + (size_t)b * 3 * P + o;
                ro[0] = g.res.x, ro[P] = g.res.y, ro[2 * P] = g.res.z;
              }
              if (MODE == 2) {   // c_consistency (Trainer.py:384-386); at level 0 up(mask) == mask, up(disp) == disp
                const float valid = __ldg(disp + o) > d.mask_disp_thrd ? 1.f : 0.f;
                s_cc[f] += valid * (1.f - m) * (fabsf(g.res.x) + fabsf(g.res.y) + fabsf(g.res.z));
              }
            } else {             // levels > 0: bilinear down-sampling = mean of the 2x2 centre taps of each 2^s block
              const int half = 1 << (shift - 1), msk = (1 << shift) - 1;
              const int rr = pass & msk & 3, cr = lane & msk;   // row / column inside the block (rows: 4w + pass, r0 % 32 == 0)
              const bool row_tap = shift == 3 ? ((4 * warp + pass) & 7) == 3 || ((4 * warp + pass) & 7) == 4 : (rr == half - 1 || rr == half);
              const bool col_tap = cr == half - 1 || cr == half;
              if (row_tap) {   // warp-uniform
                // centre columns pair up as lanes (half-1, half) of the block: xor 1 (2x2), 3 (4x4), 7 (8x8); the writer lane adds
                // the block's second centre row onto the first with a plain read-modify-write (same thread owns both rows; at
                // level 3 the two rows belong to neighbouring warps and go to two partial planes that stage C adds)
                const bool first_row = shift == 1 ? (pass & 1) == 0 : (shift == 2 ? pass == 1 : true);
                const int tl = TILE >> shift;
                float* lo = smem_lo(smem) + f * 5 * LO_PLANE + ((4 * warp + pass) >> shift) * tl + (lane >> shift) +
                            (shift == 3 && (warp & 1) ? LO_PLANE / 2 : 0);
                const float v[5] = {g.res.x, g.res.y, g.res.z, g.dsx, g.dsy};
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                  const float mine = col_tap ? v[q] : 0.f;
                  const float pair = 0.25f * (mine + __shfl_xor_sync(0xffffffffu, mine, msk));
                  if (cr == half - 1) lo[q * LO_PLANE] = first_row ? pair : lo[q * LO_PLANE] + pair;
                }
              }
            }
          }
          if (a.has_aux) {
            if (a.aux.warped[si][f]) {
              float* wout = a.aux.warped[si][f] + (size_t)b * 3 * P + o;
              wout[0] = col[0], wout[P] = col[1], wout[2 * P] = col[2];
            }
            if (a.aux.sample[si][f])
              reinterpret_cast<float2*>(a.aux.sample[si][f])[(size_t)b * P + o] = make_float2(g.gx, g.gy);
            if (MODE >= 1 && a.aux.independ[si][f]) {
              float* io = a.aux.independ[si][f] + (size_t)b * 3 * P + o;
              io[0] = g.res.x * m, io[P] = g.res.y * m, io[2 * P] = g.res.z * m;   // Trainer.py:253
            }
          }
        }
      }
    }
    __syncthreads();
    if (si + 1 < d.num_scales) stage_patch(si + 1);   // next level's patch streams in under stages B and C

    // ---- stage B: SSIM + L1 + min selection ---------------------------------------------------
    float s_photo = 0.f, s_ident = 0.f;
    {
      SsimOut wl;
      ssim_l1_run<F>(Y, smem_XI(smem), lane, warp * 4, d.ssim_weight, l1_w, wl);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + warp * 4 + k, c = c0 + lane;
        const size_t o = (size_t)r * W + c;
        float best = 0.f;
        int arg = 0;
        bool first = true;
        if (automask) {   // candidates: identity frames first (Trainer.py:339-347)
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float v = ident.L[f][k];
            if (d.noise[si]) v += __ldg(d.noise[si] + ((size_t)b * F + f) * P + o) * 0.00001f;
            if (first || v < best) best = v, arg = f, first = false;
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float v = wl.L[f][k];
          if (first || v < best) best = v, arg = (automask ? F : 0) + f, first = false;
        }
        s_photo += best;
        const float sel = (automask && arg > F - 1) ? 1.f : 0.f;   // Trainer.py:350
        s_ident += sel;
        if (automask && a.has_aux && a.aux.ident_sel[si]) a.aux.ident_sel[si][(size_t)b * P + o] = sel;
      }
    }

    // ---- stage C: low-resolution by-products of the scene-flow phases (levels > 0) -----------------
    if (MODE >= 1 && shift != 0) {
      const int tl = TILE >> shift;
      for (int i = tid; i < tl * tl; i += WP_THREADS) {
        const int li = i / tl, lj = i - li * tl;
        const int gi = (r0 >> shift) + li, gj = (c0 >> shift) + lj;
        const size_t ol = (size_t)gi * w + gj;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float* lo = smem_lo(smem) + f * 5 * LO_PLANE + i;
          float acc[5];
#pragma unroll
          for (int k = 0; k < 5; ++k)   // (level 3: the two centre rows of a block come from two warps)
            acc[k] = lo[k * LO_PLANE] + (shift == 3 ? lo[k * LO_PLANE + LO_PLANE / 2] : 0.f);
          const float mag = acc[3] * acc[3] + acc[4] * acc[4];          // Trainer.py:396
          s_mag[f] += mag;
          if (a.has_aux && a.aux.mag[si][f]) a.aux.mag[si][f][(size_t)b * p_lo + ol] = mag;
          if (a.has_aux && a.aux.resid[si][f]) {
            float* ro = a.aux.resid[si][f] + (size_t)b * 3 * p_lo + ol;
            ro[0] = acc[0], ro[p_lo] = acc[1], ro[2 * p_lo] = acc[2];
          }
          if (MODE == 2) {   // c_consistency (Trainer.py:384-386)
            const float valid = __ldg(disp + ol) > d.mask_disp_thrd ? 1.f : 0.f;
            const float ms = __ldg(d.mask[si][f] + (size_t)b * p_lo + ol);
            s_cc[f] += valid * (1.f - ms) * (fabsf(acc[0]) + fabsf(acc[1]) + fabsf(acc[2]));
          }
        }
      }
    }

    // ---- block reduction of this level's sums -------------------------------------------------
    float vals[6] = {s_photo, s_cc[0], s_cc[1], s_mag[0], s_mag[1], s_ident};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float v = warp_sum(vals[k]);
      if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();   // also guards the X tiles / side buffers before the next level overwrites them
    if (tid < 6) {
      float v = 0.f;
#pragma unroll
      for (int wi = 0; wi < WP_THREADS / 32; ++wi) v += red[wi][tid];
      a.partial[(size_t)(si * DD_NSUM + tid) * num_ctas + cta] = v;
    }
  }
}

// Deterministic second stage: sums[k] = sum over CTAs of partial[k][cta] (double accumulation).
__global__ void finalize_sums_kernel(const float* __restrict__ partial, float* __restrict__ sums, int num_ctas,
                                     int nsum_used) {
  __shared__ double sh[256 / 32];
  const int k = blockIdx.x;   // scale*DD_NSUM + slot
  double acc = 0.0;
  if ((k % DD_NSUM) < nsum_used)
    for (int i = threadIdx.x; i < num_ctas; i += blockDim.x) acc += (double)partial[(size_t)k * num_ctas + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    sums[k] = (float)t;
  }
}

int validate_desc(const dd_warp_desc* d);   // warp_photo_api.cu

template <int MODE, int F>
static int launch_fwd(const FwdArgs& args, dim3 grid, size_t smem_bytes, cudaStream_t st) {
  auto kern = warp_photo_fwd_kernel<MODE, F>;
  DD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  kern<<<grid, WP_THREADS, smem_bytes, st>>>(args); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

int warp_photo_fwd_impl(const dd_warp_desc* desc, const dd_warp_aux* aux, float* sums, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
  int rc = validate_desc(desc);
  if (rc != DD_OK) return rc;
  DD_REQUIRE(sums != nullptr, "dd_warp_photo_fwd: sums is NULL");
  const size_t need = dd_warp_photo_workspace_bytes(desc);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("dd_warp_photo_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DD_ERR_WORKSPACE;
  }
  FwdArgs args;
  args.d = *desc;
  args.has_aux = aux != nullptr;
  if (aux) args.aux = *aux; else memset(&args.aux, 0, sizeof(args.aux));
  args.partial = reinterpret_cast<float*>(workspace);
  args.min_disp = 1.f / desc->max_depth;
  args.disp_range = 1.f / desc->min_depth - 1.f / desc->max_depth;
  for (int f = 0; f < DD_MAX_FRAMES; ++f) args.tex[f] = f < desc->num_frames ? source_texture(desc->source[f], desc->B, desc->H, desc->W) : 0;
  const dim3 grid(desc->W / TILE, desc->H / TILE, desc->B);
  const int num_ctas = grid.x * grid.y * grid.z;
  const int mode = (desc->flags & DD_FLAG_CMPFLOW) ? ((desc->flags & DD_FLAG_MOTMASK) ? 2 : 1) : 0;
  size_t smem_bytes = 9 * PLANE1 * sizeof(float);
  if (mode >= 1) smem_bytes += 2 * 5 * LO_PLANE * sizeof(float);
  smem_bytes += (mode == 0 ? 1 : (mode == 1 ? 1 + 3 * desc->num_frames : 1 + 4 * desc->num_frames)) * PATCH_PLANE * sizeof(float);
  const int F = desc->num_frames;
  if (mode == 0 && F == 2) rc = launch_fwd<0, 2>(args, grid, smem_bytes, st);
  else if (mode == 1 && F == 2) rc = launch_fwd<1, 2>(args, grid, smem_bytes, st);
  else if (mode == 2 && F == 2) rc = launch_fwd<2, 2>(args, grid, smem_bytes, st);
  else if (mode == 0 && F == 1) rc = launch_fwd<0, 1>(args, grid, smem_bytes, st);
  else if (mode == 1 && F == 1) rc = launch_fwd<1, 1>(args, grid, smem_bytes, st);
  else rc = launch_fwd<2, 1>(args, grid, smem_bytes, st);
  if (rc != DD_OK) return rc;
  finalize_sums_kernel<<<desc->num_scales * DD_NSUM, 256, 0, st>>>(args.partial, sums, num_ctas, 6); dd::count_launches(1);
  DD_CHECK_CUDA(cudaGetLastError());
  return DD_OK;
}

}  // namespace dd
