// Per-pixel geometry of the view-synthesis path, shared by the forward and backward kernels.
// Arithmetic follows the reference operation by operation (see each comment) so that fp32
// rounding stays as close to the PyTorch path as an independent implementation can be.
#pragma once
#include "dd_common.cuh"

namespace dd {

constexpr int TILE = 32;            // CTA output tile (H, W are multiples of 32: Trainer.py:25-26)
constexpr int WP_THREADS = 256;

// per-image camera constants staged in shared memory
struct CamConst {
  float K[12];      // K[:3,:]            (tools.py:214)
  float iK[9];      // inv_K[:3,:3]       (tools.py:193)
  float T[2][16];   // cam_T_cam per source frame
  float ts[2];      // ('ts', f)
  float inv_wm1, inv_hm1;   // 1/(W-1), 1/(H-1) of Project3D's normalisation (tools.py:219-220)
};

__device__ __forceinline__ void load_cam(CamConst* cam, const dd_warp_desc& d, int b, int tid) {
  if (tid < 12) cam->K[tid] = __ldg(d.K + b * 16 + tid);
  if (tid >= 32 && tid < 41) {
    const int i = tid - 32;
    cam->iK[i] = __ldg(d.inv_K + b * 16 + (i / 3) * 4 + (i % 3));
  }
  if (tid >= 64 && tid < 64 + 16 * d.num_frames) {
    const int i = tid - 64;
    cam->T[i / 16][i % 16] = __ldg(d.T[i / 16] + b * 16 + (i % 16));
  }
  if (tid >= 128 && tid < 128 + d.num_frames) {
    const int f = tid - 128;
    cam->ts[f] = d.ts[f] ? __ldg(d.ts[f] + b) : 1.f;
  }
  if (tid == 160) {
    cam->inv_wm1 = 1.f / (float)(d.W - 1);
    cam->inv_hm1 = 1.f / (float)(d.H - 1);
  }
}

struct Vec3 {
  float x, y, z;
};
struct Vec4 {
  float x, y, z, w;
};

// X = T @ (p, 1)   (tools.py:213)
__device__ __forceinline__ Vec4 apply_T(const float* T, const Vec3& p) {
  Vec4 X;
  X.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
  X.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
  X.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
  X.w = T[12] * p.x + T[13] * p.y + T[14] * p.z + T[15];
  return X;
}

// c = K[:3,:] @ X ; pix = c[:2] / (c[2] + eps)   (tools.py:214-216)
struct Proj {
  float c0, c1, z;   // z = c2 + eps
  float iz;          // 1 / z
  float px, py;      // pixel coordinates
};
__device__ __forceinline__ Proj project_K(const float* K, const Vec4& X) {
  Proj p;
  p.c0 = K[0] * X.x + K[1] * X.y + K[2] * X.z + K[3] * X.w;
  p.c1 = K[4] * X.x + K[5] * X.y + K[6] * X.z + K[7] * X.w;
  p.z = (K[8] * X.x + K[9] * X.y + K[10] * X.z + K[11] * X.w) + 1e-7f;
  // one correctly rounded reciprocal + two multiplies instead of two divisions (<= 1.5 ulp from c/z)
  p.iz = rcp_nr(p.z);
  p.px = p.c0 * p.iz;
  p.py = p.c1 * p.iz;
  return p;
}

// normalised grid coordinate as Project3D stores it: (pix/(size-1) - 0.5) * 2   (tools.py:219-221)
__device__ __forceinline__ float normalise(float pix, float inv_size_m1) { return (pix * inv_size_m1 - 0.5f) * 2.f; }
// grid_sample(align_corners=True) un-normalisation: ((g+1)/2) * (size-1)   (GridSampler.cuh:23-30)
__device__ __forceinline__ float unnormalise(float g, int size) { return ((g + 1.f) * 0.5f) * (float)(size - 1); }

// Border-clamped bilinear sampling footprint (padding_mode='border', GridSampler.cuh:55-57)
struct Foot {
  int x0, y0, x1, y1;     // x1/y1 clamped into the image (their weight is 0 when they were outside)
  float wx0, wx1, wy0, wy1;  // (x1-ix), (ix-x0), (y1-iy), (iy-y0)
  bool live_x, live_y;    // coordinate gradient passes (strictly inside (0, size-1))
};
__device__ __forceinline__ Foot footprint(float ix, float iy, int H, int W) {
  Foot ft;
  ft.live_x = (ix > 0.f) && (ix < (float)(W - 1));
  ft.live_y = (iy > 0.f) && (iy < (float)(H - 1));
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float fx = floorf(ix), fy = floorf(iy);
  ft.x0 = (int)fx;
  ft.y0 = (int)fy;
  ft.wx1 = ix - fx;
  ft.wx0 = (fx + 1.f) - ix;
  ft.wy1 = iy - fy;
  ft.wy0 = (fy + 1.f) - iy;
  ft.x1 = min(ft.x0 + 1, W - 1);
  ft.y1 = min(ft.y0 + 1, H - 1);
  return ft;
}

// The four texels of a sampling footprint.  With a texture object over the source tensor (all B*3 planes stacked as one
// pitch-linear 2-D image of B*3*H rows, point filtering, clamp addressing) ONE tex2Dgather instruction returns them
// exactly (no hardware interpolation: the weights stay fp32 in the kernel), through the texture pipe instead of four
// address computations + four LSU gathers.  The gather is addressed at the centre of the 2x2 block, (x0 + 1, row + y0 + 1),
// so which texels are returned does not depend on any rounding of the coordinate.  A texel beyond the right / bottom border
// of a plane is replaced by whatever the clamp / the next plane holds -- its bilinear weight is exactly 0 in that case
// (footprint(): x0 = W - 1 only for ix = W - 1).  tex == 0: plain loads.
struct Quad {
  float nw, ne, sw, se;
};
__device__ __forceinline__ Quad gather4(cudaTextureObject_t tex, const float* __restrict__ img, int W, int row_base, const Foot& ft) {
  Quad q;
  if (tex != 0) {
    const float4 t = tex2Dgather<float4>(tex, (float)ft.x0 + 1.f, (float)(row_base + ft.y0) + 1.f, 0);
    q.sw = t.x, q.se = t.y, q.ne = t.z, q.nw = t.w;   // (i, j+1), (i+1, j+1), (i+1, j), (i, j)
  } else {
    q.nw = __ldg(img + ft.y0 * W + ft.x0), q.ne = __ldg(img + ft.y0 * W + ft.x1);
    q.sw = __ldg(img + ft.y1 * W + ft.x0), q.se = __ldg(img + ft.y1 * W + ft.x1);
  }
  return q;
}
__device__ __forceinline__ float blend(const Quad& q, const Foot& ft) {
  return q.nw * (ft.wx0 * ft.wy0) + q.ne * (ft.wx1 * ft.wy0) + q.sw * (ft.wx0 * ft.wy1) + q.se * (ft.wx1 * ft.wy1);
}
__device__ __forceinline__ float sample_plane(cudaTextureObject_t tex, const float* __restrict__ img, int W, int row_base, const Foot& ft) {
  return blend(gather4(tex, img, W, row_base, ft), ft);
}

// host side (api.cu): cached texture object over a (B,3,H,W) fp32 tensor, 0 when the tensor does not qualify
cudaTextureObject_t source_texture(const float* ptr, int B, int H, int W);

// Low-resolution operands (disp_s, flow_s, mask_s at levels > 0) are staged per tile in shared memory with 4-byte
// cp.async copies: ((tile >> s) + 2)^2 texels per plane cover every tap the tile's pixels interpolate from.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::); }

// bilinear up-sampling from a staged patch (same expression as bilerp(): identical rounding)
__device__ __forceinline__ float patch_bilerp(const float* __restrict__ pl, int pw, int y0, int y1, int x0, int x1, float ly, float lx) {
  const float v00 = pl[y0 * pw + x0], v01 = pl[y0 * pw + x1], v10 = pl[y1 * pw + x0], v11 = pl[y1 * pw + x1];
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// Everything one pixel of one (scale, frame) needs.  MODE: 0 rigid, 1 CmpFlow, 2 CmpFlow+MotMask
// (Trainer.py:248-278).
struct PixelGeom {
  Vec3 ray;      // inv_K[:3,:3] @ (u,v,1)
  float depth;   // 1 / (min_disp + range*disp_up)
  Vec3 Pc;       // depth * ray
};

struct FrameGeom {
  Vec3 Pin;      // the 3-vector fed (with w=1) to T (MODE 0/2) or directly to K (MODE 1)
  Vec4 X;        // camera-frame point handed to K
  Proj pr;       // projection of X
  float gx, gy;  // normalised sample coordinate ('sample', f, s)
  // flow modes only
  Vec3 cf;       // up(complete_flow_s) * ts
  Vec3 res;      // residual flow  cf - ego
  float m;       // up(motion_mask_s) (1 in MODE 1)
  float dsx, dsy;  // sample_ego - sample_complete (normalised units)
};

template <int MODE>
__device__ __forceinline__ void frame_geometry(FrameGeom& g, const PixelGeom& pg, const CamConst* cam, int f,
                                               const Vec3& cf_up, float m_up, int H, int W, bool want_side) {
  const float* T = cam->T[f];
  if (MODE == 0) {
    g.Pin = pg.Pc;
    g.X = apply_T(T, pg.Pc);
  } else {
    const Vec4 Xe = apply_T(T, pg.Pc);                       // Trainer.py:250
    const Vec3 ego = {Xe.x - pg.Pc.x, Xe.y - pg.Pc.y, Xe.z - pg.Pc.z};
    g.cf = cf_up;
    g.res = {cf_up.x - ego.x, cf_up.y - ego.y, cf_up.z - ego.z};   // Trainer.py:252
    g.m = m_up;
    if (want_side) {
      const Proj pe = project_K(cam->K, Xe);
      const Vec4 Xc = {pg.Pc.x + cf_up.x, pg.Pc.y + cf_up.y, pg.Pc.z + cf_up.z, 1.f};   // Trainer.py:257-260
      const Proj pc = project_K(cam->K, Xc);
      g.dsx = normalise(pe.px, cam->inv_wm1) - normalise(pc.px, cam->inv_wm1);
      g.dsy = normalise(pe.py, cam->inv_hm1) - normalise(pc.py, cam->inv_hm1);
    }
    if (MODE == 2) {
      g.Pin = {pg.Pc.x + g.res.x * m_up, pg.Pc.y + g.res.y * m_up, pg.Pc.z + g.res.z * m_up};   // Trainer.py:265-267
      g.X = apply_T(T, g.Pin);
    } else {
      g.Pin = {pg.Pc.x + cf_up.x, pg.Pc.y + cf_up.y, pg.Pc.z + cf_up.z};                         // Trainer.py:270-271
      g.X = {g.Pin.x, g.Pin.y, g.Pin.z, 1.f};
    }
  }
  g.pr = project_K(cam->K, g.X);
  g.gx = normalise(g.pr.px, cam->inv_wm1);
  g.gy = normalise(g.pr.py, cam->inv_hm1);
}

}  // namespace dd
