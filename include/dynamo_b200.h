/*
 * dynamo_b200.h -- C ABI of libdynamo_b200.so (hand-written sm_100a CUDA kernels for the
 * per-step hot path of Dynamo-Depth).
 *
 * The reference (YihongSun/Dynamo-Depth @227a5d9) is pure Python/PyTorch and has no FFI of its own;
 * its boundary for this path is a set of Python callables.  Every entry point
 * below names the reference code it replaces (file:line relative to the reference root).  The
 * Python host side (dynamo-depth_b200/dd_b200/) binds these with ctypes and keeps the reference's
 * module surface (tools.py, networks/layers.py, Trainer.generate_images_pred / compute_losses).
 *
 * Conventions
 *   - all tensors are fp32, contiguous, NCHW, resident in device memory; pointers are raw
 *     device addresses; sizes are plain ints.  No torch types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream,
 *     never synchronises, never allocates.  Scratch memory comes from the caller (`workspace`),
 *     sized by the matching dd_*_workspace_bytes().
 *   - return value: 0 on success, negative on error; dd_last_error() returns a thread-local
 *     message.  There is no CPU fallback anywhere: a missing device is an error.
 */
#ifndef DYNAMO_B200_H
#define DYNAMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_OK 0
#define DD_ERR_INVALID (-1)
#define DD_ERR_CUDA (-2)
#define DD_ERR_WORKSPACE (-3)

#define DD_MAX_SCALES 4
#define DD_MAX_FRAMES 2

/* flags of dd_warp_desc */
#define DD_FLAG_CMPFLOW 1  /* Model.bool_CmpFlow: complete 3-D flow is predicted (Trainer.py:248)   */
#define DD_FLAG_MOTMASK 2  /* Model.bool_MotMask: motion mask gates the residual flow (Trainer.py:262) */
#define DD_FLAG_AUTOMASK 4 /* Trainer.bool_automask: identity reprojection candidates (Trainer.py:327) */

/* per-scale sums produced by dd_warp_photo_fwd (slot index inside sums[scale][DD_NSUM]) */
#define DD_NSUM 8
#define DD_SUM_PHOTO 0     /* sum over B,H,W of min-selected reprojection loss (Trainer.py:347-352) */
#define DD_SUM_CONSIST0 1  /* +f: sum over B,3,h,w of valid*(1-mask)*|residual_flow_s| (Trainer.py:385-386) */
#define DD_SUM_MAG0 3      /* +f: sum over B,h,w of ||down(sample_ego)-down(sample_complete)||^2 (Trainer.py:394-397) */
#define DD_SUM_IDENT 5     /* number of pixels whose argmin is a warped (non-identity) candidate */

const char* dd_last_error(void);
int dd_version(void);
/* number of SMs / device index the library sees (negative on error) */
int dd_device_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Fused view synthesis + photometric loss
 *   replaces Trainer.generate_images_pred (Trainer.py:215-287) and the photometric / automask /
 *   c_consistency / disp_mag parts of Trainer.compute_losses (Trainer.py:312-352,384-397), i.e.
 *   utils.interp (utils.py:98-101), tools.disp_to_depth (tools.py:291-298), tools.BackprojectDepth
 *   (tools.py:167-197), tools.Project3D (tools.py:200-224), F.grid_sample (Trainer.py:281),
 *   tools.SSIM (tools.py:227-257) and Trainer.compute_reprojection_loss (Trainer.py:413-423),
 *   for all pyramid levels and both source frames in one launch.
 * ------------------------------------------------------------------------------------------ */
typedef struct dd_warp_desc {
  int32_t B, H, W;            /* batch, full resolution (H, W multiples of 32: Trainer.py:25-26) */
  int32_t num_scales;         /* 1..DD_MAX_SCALES pyramid levels handled by this call */
  int32_t num_frames;         /* 1..DD_MAX_FRAMES source frames (opt.frame_ids[1:]) */
  int32_t flags;              /* DD_FLAG_* */
  float min_depth, max_depth; /* options.py:182-189 */
  float ssim_weight;          /* options.py:115-118 */
  float mask_disp_thrd;       /* options.py:119-122 */
  const float* target;                /* ('color',0,0)            (B,3,H,W) */
  const float* source[DD_MAX_FRAMES]; /* ('color',f,0)            (B,3,H,W) */
  const float* K;                     /* ('K',0)                  (B,4,4)   */
  const float* inv_K;                 /* ('inv_K',0)              (B,4,4)   */
  const float* T[DD_MAX_FRAMES];      /* ('cam_T_cam',0,f)        (B,4,4)   */
  const float* ts[DD_MAX_FRAMES];     /* ('ts',f) as fp32 (B,) or NULL (=1) */
  int32_t scale[DD_MAX_SCALES];       /* s: level resolution is (H>>s, W>>s) */
  const float* disp[DD_MAX_SCALES];   /* ('disp',0,s)             (B,1,h,w) */
  const float* flow[DD_MAX_SCALES][DD_MAX_FRAMES]; /* ('complete_flow',f,s) (B,3,h,w) [CMPFLOW] */
  const float* mask[DD_MAX_SCALES][DD_MAX_FRAMES]; /* ('motion_mask',f,s)   (B,1,h,w) [MOTMASK] */
  const float* noise[DD_MAX_SCALES];  /* automask tie-break N(0,1) (B,F,H,W) or NULL (Trainer.py:339) */
} dd_warp_desc;

/* optional materialised by-products of the forward pass (any pointer may be NULL) */
typedef struct dd_warp_aux {
  float* warped[DD_MAX_SCALES][DD_MAX_FRAMES];   /* ('color',f,s)            (B,3,H,W) */
  float* sample[DD_MAX_SCALES][DD_MAX_FRAMES];   /* ('sample',f,s)           (B,H,W,2) normalised */
  float* depth[DD_MAX_SCALES];                   /* ('depth',0,s)            (B,1,H,W) */
  float* ident_sel[DD_MAX_SCALES];               /* 'identity_selection/s'   (B,H,W)   */
  float* resid[DD_MAX_SCALES][DD_MAX_FRAMES];    /* ('residual_flow',f,s)    (B,3,h,w) */
  float* independ[DD_MAX_SCALES][DD_MAX_FRAMES]; /* ('independ_flow',f,s)    (B,3,H,W) */
  float* mag[DD_MAX_SCALES][DD_MAX_FRAMES];      /* ||down(sample_ego-sample_complete)||^2 (B,h,w) */
} dd_warp_aux;

size_t dd_warp_photo_workspace_bytes(const dd_warp_desc* desc);

/* sums: device (num_scales, DD_NSUM) fp32, overwritten.  aux may be NULL. */
int dd_warp_photo_fwd(const dd_warp_desc* desc, const dd_warp_aux* aux, float* sums,
                      void* workspace, size_t workspace_bytes, void* stream);

/* gradients of  L = sum_{s,k} grad_sums[s][k] * sums[s][k]  (k in PHOTO, CONSIST0+f)
 * w.r.t. disp / cam_T_cam / complete_flow / motion_mask.  Output buffers are overwritten
 * (pointers that are NULL are skipped).  grad_T[f] (B,4,4) accumulates over all scales. */
typedef struct dd_warp_grads {
  float* disp[DD_MAX_SCALES];                 /* (B,1,h,w) */
  float* T[DD_MAX_FRAMES];                    /* (B,4,4)   */
  float* flow[DD_MAX_SCALES][DD_MAX_FRAMES];  /* (B,3,h,w) */
  float* mask[DD_MAX_SCALES][DD_MAX_FRAMES];  /* (B,1,h,w); the two frames must not alias */
} dd_warp_grads;

/* saved: by-products of the matching forward call; saved->resid[s][f] is required for levels with
 * scale[s] > 0 when DD_FLAG_MOTMASK is set (sign of the down-sampled residual flow), else may be NULL. */
int dd_warp_photo_bwd(const dd_warp_desc* desc, const float* grad_sums, const dd_warp_aux* saved,
                      const dd_warp_grads* grads, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAMO_B200_H */
