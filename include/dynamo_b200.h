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
/* number of kernels this library has launched in this process so far (bookkeeping for benchmarks) */
long long dd_launch_count(void);

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

/* ------------------------------------------------------------------------------------------
 * Edge-aware smoothness, batched over every (term, level) of a step
 *   replaces tools.compute_smooth_loss (tools.py:311-326) and the mean-normalised disparity of
 *   Trainer.py:357-358:  sum_x = sum |n[..,:-1]-n[..,1:]| * exp(-mean_c|img[..,:-1]-img[..,1:]|),
 *   sum_y likewise along rows, n = inp / (mean_hw(inp) + 1e-7) when mean_normalise is set.
 *   The caller divides by the element counts B*C*h*(w-1) and B*C*(h-1)*w (two separate means).
 * ------------------------------------------------------------------------------------------ */
#define DD_MAX_SMOOTH_TASKS 24
typedef struct dd_smooth_task {
  const float* inp;       /* (B,C,h,w) */
  const float* img;       /* (B,3,h,w) colour of the same level, or NULL (no edge weights) */
  float* grad_inp;        /* backward output (B,C,h,w), overwritten; NULL = skip this task */
  int32_t B, C, h, w;
  int32_t mean_normalise; /* Trainer.py:357-358 */
} dd_smooth_task;

size_t dd_smooth_workspace_bytes(const dd_smooth_task* tasks, int ntasks);
/* sums: device (ntasks, 2) = (sum_x, sum_y) */
int dd_smooth_fwd(const dd_smooth_task* tasks, int ntasks, float* sums, void* workspace, size_t workspace_bytes,
                  void* stream);
/* grad_sums: device (ntasks, 2) upstream gradients of (sum_x, sum_y) */
int dd_smooth_bwd(const dd_smooth_task* tasks, int ntasks, const float* grad_sums, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Motion-mask sparsity (Trainer.py:388-399)
 *   static = mag < mean(mag)   (mag = ||down(sample_ego)-down(sample_complete)||^2 from
 *   dd_warp_photo_fwd, mean over the whole batch), loss = mean over static pixels of
 *   BCEWithLogits(prob, 0) = softplus(prob), skipped (0) unless every image has a static pixel.
 *   The reference's two host synchronisations (torch.all / boolean indexing) become a
 *   device-side guard.
 *   out[0] = loss, out[1] = guard / count (backward scale), out[2] = guard, out[3] = count
 * ------------------------------------------------------------------------------------------ */
size_t dd_msparsity_workspace_bytes(int B, int h, int w);
int dd_msparsity_fwd(const float* mag, const float* mag_sum, const float* prob, int B, int h, int w, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);
/* grad_prob (B,1,h,w) = grad_out[0] * out[1] * static * sigmoid(prob), overwritten */
int dd_msparsity_bwd(const float* mag, const float* mag_sum, const float* prob, const float* out,
                     const float* grad_out, int B, int h, int w, float* grad_prob, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused decoder convolution (fp32 SIMT implicit GEMM; tensor cores are reserved for the encoders)
 *   replaces networks/layers.py ConvBlock / Conv3x3 / upsample (layers.py:85-121) as used by
 *   DepthDecoder / LiteDepthDecoder (depth_decoder.py:40-55,99-115), the refine convolutions of
 *   MotionDecoder (motion_decoder.py:24-62) and the PoseDecoder convolutions (pose_decoder.py:16-37):
 *     out = act( conv_k( pad( concat( up(x0), x1 ) ) ) + bias ) [+ residual]
 *   with pad = ReflectionPad2d(1) | zero, up = none | nearest x2 | bilinear x2 (align_corners=False),
 *   act = none | ELU | sigmoid | ReLU, all in one pass over the inputs.
 * ------------------------------------------------------------------------------------------ */
#define DD_PAD_ZERO 0
#define DD_PAD_REFLECT 1
#define DD_UP_NONE 0
#define DD_UP_NEAREST2 1
#define DD_UP_BILINEAR2 2
#define DD_ACT_NONE 0
#define DD_ACT_ELU 1
#define DD_ACT_SIGMOID 2
#define DD_ACT_RELU 3

typedef struct dd_conv_desc {
  int32_t B, H, W;       /* output (= convolution input) spatial size */
  int32_t Cout;
  int32_t ksize;         /* 1 or 3 (stride 1, "same" padding) */
  int32_t pad_mode;      /* DD_PAD_* (ignored for ksize 1) */
  int32_t act;           /* DD_ACT_* */
  int32_t up0;           /* DD_UP_*: how x0 reaches (H, W) */
  int32_t C0, C1;        /* channels of x0 and of the optional skip tensor x1 (0 = none) */
  const float* x0;       /* (B,C0,H,W) or (B,C0,H/2,W/2) when up0 != DD_UP_NONE */
  const float* x1;       /* (B,C1,H,W) or NULL */
  const float* weight;   /* (Cout, C0+C1, k, k) OIHW as in nn.Conv2d */
  const float* bias;     /* (Cout) or NULL */
  const float* residual; /* (B,Cout,H,W) added after the activation, or NULL */
} dd_conv_desc;

size_t dd_conv_workspace_bytes(const dd_conv_desc* desc);
int dd_conv_fwd(const dd_conv_desc* desc, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* out: forward result (needed for ELU / sigmoid / ReLU derivatives, may be NULL for DD_ACT_NONE);
 * grad_x0 / grad_x1 / grad_weight / grad_bias may each be NULL (skipped); all are overwritten.
 * The gradient w.r.t. `residual` is grad_out itself. */
int dd_conv_bwd(const dd_conv_desc* desc, const float* out, const float* grad_out, float* grad_x0, float* grad_x1,
                float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bilinear resize, align_corners=False (utils.interp, utils.py:98-101; F.interpolate at
 * motion_decoder.py:38 and depth_decoder.py:112), optional fused sigmoid (depth_decoder.py:113).
 * ------------------------------------------------------------------------------------------ */
int dd_resize_bilinear_fwd(const float* x, int BC, int h_in, int w_in, int h_out, int w_out, int sigmoid, float* out,
                           void* stream);
/* grad_x (BC,h_in,w_in) overwritten; `out` is the forward result (needed when sigmoid != 0) */
int dd_resize_bilinear_bwd(const float* grad_out, const float* out, int BC, int h_in, int w_in, int h_out, int w_out,
                           int sigmoid, float* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input colour pyramid: out (BC, H/2, W/2) = clamp(bicubic-antialias x1/2 of x (BC, H, W), 0, 1) —
 * one link of the chain Trainer.apply_img_resize builds (Trainer.py:729-734 with the torchvision
 * Resize(BICUBIC, antialias=True) of Trainer.py:80, i.e. ATen _upsample_bicubic2d_aa).  H, W even.  No gradient
 * (the pyramid is input data: it only feeds the edge weights of compute_smooth_loss, tools.py:311-326).
 * ------------------------------------------------------------------------------------------ */
int dd_pyramid_half_fwd(const float* x, int BC, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone geometry / photometric layers (the module surface eval scripts and user code call:
 * trainer.backproject_depth[s](depth, inv_K), trainer.project_3d[s](points, K, T), SSIM()(x, y)).
 * The training step itself goes through dd_warp_photo_* and never materialises these tensors.
 * ------------------------------------------------------------------------------------------ */
/* tools.BackprojectDepth.forward (tools.py:191-197): points (B,4,H*W) = cat(depth * inv_K[:3,:3] @ (u,v,1), 1) */
int dd_backproject_fwd(const float* depth, const float* inv_K, int B, int H, int W, float* points, void* stream);
int dd_backproject_bwd(const float* grad_points, const float* inv_K, int B, int H, int W, float* grad_depth, void* stream);
/* tools.Project3D.forward (tools.py:211-224): pix (B,H,W,2) normalised to [-1,1], ego (B,3,H*W); T may be NULL */
int dd_project_fwd(const float* points, const float* K, const float* T, int B, int H, int W, float* pix, float* ego,
                   void* stream);
/* grad_pix / grad_ego may be NULL (treated as zero); grad_points (B,4,H*W) and grad_T (B,4,4, may be NULL) overwritten */
int dd_project_bwd(const float* points, const float* K, const float* T, const float* grad_pix, const float* grad_ego, int B,
                   int H, int W, float* grad_points, float* grad_T, void* stream);
/* tools.SSIM.forward (tools.py:243-257): out (B,C,H,W) = clamp((1 - SSIM(x,y)) / 2, 0, 1) */
int dd_ssim_fwd(const float* x, const float* y, int BC, int H, int W, float* out, void* stream);
/* gradient w.r.t. x (SSIM is symmetric: swap x and y for the gradient w.r.t. y) */
int dd_ssim_bwd(const float* x, const float* y, const float* grad_out, int BC, int H, int W, float* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pose head epilogue (pose_decoder.py:39-44 and networks/layers.py:7-82)
 *   dd_pose_mean:   out (B,C) = scale * mean over (h,w) of x (B,C,h,w)           [0.01 * out.mean(3).mean(2)]
 *   dd_pose_matrix: T (B,4,4) from axis-angle (B,3) and translation (B,3): Rodrigues with axis = v/(|v|+1e-7),
 *                   invert != 0 -> R^T @ Trans(-t), else Trans(t) @ R  (transformation_from_parameters)
 * ------------------------------------------------------------------------------------------ */
int dd_pose_mean_fwd(const float* x, int BC, int hw, float scale, float* out, void* stream);
int dd_pose_mean_bwd(const float* grad_out, int BC, int hw, float scale, float* grad_x, void* stream);
int dd_pose_matrix_fwd(const float* axisangle, const float* translation, int B, int invert, float* T, void* stream);
int dd_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_T, int B, int invert,
                       float* grad_axisangle, float* grad_translation, void* stream);

/* ------------------------------------------------------------------------------------------
 * RANSAC ground-plane hypothesis scoring (tools.GroundPlane.estimate_ground_plane, tools.py:113-139)
 *   counts[k] = #{ n : | x_n*w[k][0] + z_n*w[k][1] + w[k][2] - y_n | < tol }  over the ground rows
 *   (row >= row0) of image (k % B) -- the reference pairs hypothesis k with image k % B because it tiles
 *   the points with .repeat(max_it,1,1) (tools.py:131); reproduced as is.  Reads every point once per
 *   20 hypotheses instead of materialising the (B*max_it, N, 3) tensor (1.9 GB at bs32, 192x640).
 * ------------------------------------------------------------------------------------------ */
int dd_ground_score(const float* points /* (B,3,H,W) */, const float* w /* (K,3), K = B*max_it */, int B, int H, int W,
                    int row0, int K, float tol, int32_t* counts /* (K) overwritten */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Lite-Mono encoder linear layers (nn.Linear at networks/depth_encoder.py:58-60 XCA qkv / proj, :197-199 and
 * :243-245 pwconv1 / pwconv2; called at :65,81,211-213,267-269) on the tcgen05 tensor cores at fp32 accuracy
 * (3xTF32 operand split, fp32 accumulation in TMEM; csrc/linear_tc.cu).  Row-major fp32, 16-byte aligned pointers,
 * K % 4 == 0, N % 4 == 0.
 *   dd_linear_fwd: y (M,N) = x (M,K) . w (N,K)^T + bias (N) [bias may be NULL]
 *   dd_linear_bwd: grad_x (M,K) = grad_y . w;  grad_w (N,K) = grad_y^T . x;  grad_b (N) = column sums of grad_y.
 *                  Each output may be NULL (grad_b needs grad_w: it is summed by the weight-gradient pass); all
 *                  are overwritten.
 * Workspace (dd_linear_workspace_bytes(M, K, N) bytes, 16-byte aligned): the weight operand of the forward and
 * input-gradient contractions is split once per call into the (hi, lo) shared-memory tile images the kernel's K blocks
 * fetch with one bulk copy each.
 * ------------------------------------------------------------------------------------------ */
size_t dd_linear_workspace_bytes(int M, int K, int N);
int dd_linear_fwd(const float* x, const float* w, const float* bias, int M, int K, int N, float* y, void* workspace, size_t workspace_bytes,
                  void* stream);
int dd_linear_bwd(const float* x, const float* w, const float* grad_y, int M, int K, int N, float* grad_x, float* grad_w,
                  float* grad_b, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout glue of the Lite-Mono blocks (networks/depth_encoder.py:204-221 DilatedConv.forward, :252-279 LGFI.forward;
 * csrc/lite_glue.cu): tiled transposes instead of PyTorch's strided element-wise kernels.
 *   dd_nchw_to_nhwc : out (B,HW,C) = x (B,C,HW) transposed  -- `x.permute(0, 2, 3, 1)` made contiguous (:210, :256)
 *   dd_nhwc_to_nchw : the inverse (its gradient)
 *   dd_block_tail_fwd : out (B,C,HW) = x (B,C,HW) + scale[b] * gamma[c] * y (B,HW,C)  -- layer scale, stochastic
 *                       depth (timm DropPath factor per sample) and residual add of :214-219 / :272-277;
 *                       gamma and scale may be NULL (= 1)
 *   dd_block_tail_bwd : grad_y (B,HW,C) = scale[b] * gamma[c] * grad_out (B,C,HW);
 *                       grad_gamma[c] = sum_{b,p} scale[b] * grad_out[b,c,p] * y[b,p,c]   (either may be NULL;
 *                       the gradient w.r.t. x is grad_out itself)
 * ------------------------------------------------------------------------------------------ */
int dd_nchw_to_nhwc(const float* x, int B, int C, int HW, float* out, void* stream);
int dd_nhwc_to_nchw(const float* x, int B, int C, int HW, float* out, void* stream);
int dd_block_tail_fwd(const float* x, const float* y, const float* gamma, const float* scale, int B, int C, int HW, float* out,
                      void* stream);
int dd_block_tail_bwd(const float* grad_out, const float* y, const float* gamma, const float* scale, int B, int C, int HW,
                      float* grad_y, float* grad_gamma, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode BatchNorm2d (+ exact GELU) of the Lite-Mono encoder over NCHW fp32 tensors (csrc/batchnorm.cu):
 * networks/depth_encoder.py:113-122 `BNGELU` (nn.BatchNorm2d(eps=1e-5) -> nn.GELU()) of the stem convolutions :131-145 and
 * `DilatedConv.bn1` :194,:208.  x, y, grad_y, grad_x are (B,C,HW); gamma / beta (C) may be NULL (= 1 / 0).
 *   dd_bn_gelu_fwd : batch statistics over (B,HW) per channel (biased variance), y = [gelu]((x - mean) * invstd * gamma + beta);
 *                    save_mean / save_invstd (C) are written for the backward pass; running_mean / running_var (may be NULL)
 *                    are updated in place as nn.BatchNorm2d does: r = (1 - momentum) * r + momentum * stat, with the
 *                    unbiased variance
 *   dd_bn_gelu_bwd : grad_x (may be NULL), grad_gamma, grad_beta (may be NULL) from x, grad_y and the saved statistics;
 *                    with gelu != 0 grad_y is the gradient w.r.t. the GELU output (the BN output is re-derived from x)
 * Both need dd_bn_workspace_bytes(C) bytes of workspace (per-CTA partial sums; deterministic reduction, no atomics).
 * ------------------------------------------------------------------------------------------ */
size_t dd_bn_workspace_bytes(int C);
int dd_bn_gelu_fwd(const float* x, int B, int C, int HW, const float* gamma, const float* beta, float eps, float momentum, int gelu,
                   float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* workspace,
                   size_t workspace_bytes, void* stream);
int dd_bn_gelu_bwd(const float* x, const float* grad_y, int B, int C, int HW, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int gelu, float* grad_x, float* grad_gamma, float* grad_beta,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * channels_last LayerNorm of the Lite-Mono LGFI blocks (csrc/layernorm.cu): networks/depth_encoder.py:90-104
 * `LayerNorm.forward` -> F.layer_norm over the last dimension (:261 norm_xca, :266 norm).  x, y, grad_y, grad_x are (M, C)
 * row-major, C a multiple of 4 and <= 512; gamma / beta (C) may be NULL (= 1 / 0); all pointers 16-byte aligned.
 *   dd_layernorm_fwd : y = (x - mean_row) * rstd_row * gamma + beta with the biased row variance; mean / rstd (M) are kept
 *                      for the backward pass
 *   dd_layernorm_bwd : grad_x (may be NULL), grad_gamma, grad_beta (may be NULL; they need
 *                      dd_layernorm_workspace_bytes(C) bytes of workspace: per-CTA partials, fixed-order reduction)
 * ------------------------------------------------------------------------------------------ */
size_t dd_layernorm_workspace_bytes(int C);
int dd_layernorm_fwd(const float* x, long long M, int C, const float* gamma, const float* beta, float eps, float* y, float* mean,
                     float* rstd, void* stream);
int dd_layernorm_bwd(const float* x, const float* grad_y, long long M, int C, const float* gamma, const float* mean, const float* rstd,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Depth-wise dilated 3x3 convolution of the Lite-Mono DilatedConv blocks (csrc/dwconv.cu): networks/depth_encoder.py:148-168
 * `CDilated` = nn.Conv2d(dim, dim, 3, stride=1, padding=d, dilation=d, groups=dim, bias=False) as used at :193/:207.
 * x, y, grad_y are (B,C,H,W) NCHW with W a multiple of 4, w is (C,1,3,3); dilation in {1, 2, 3, 4, 6}.
 *   dd_dwconv3x3_fwd   : y = conv(x, w); flip != 0 applies the taps mirrored, i.e. computes the data gradient when called
 *                        on grad_y
 *   dd_dwconv3x3_wgrad : grad_w (C,1,3,3) from x and grad_y; needs dd_dwconv3x3_workspace_bytes(C) bytes (per-CTA partials,
 *                        fixed-order reduction)
 * ------------------------------------------------------------------------------------------ */
size_t dd_dwconv3x3_workspace_bytes(int C);
int dd_dwconv3x3_fwd(const float* x, const float* w, int B, int C, int H, int W, int dilation, int flip, float* y, void* stream);
int dd_dwconv3x3_wgrad(const float* x, const float* grad_y, int B, int C, int H, int W, int dilation, float* grad_w, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the ResNet trunks (csrc/pool.cu): networks/resnet_encoder.py:18,:130,
 * on channels_last activations, i.e. x is (B,H,W,C) and y (B,Ho,Wo,C) in memory, Ho = (H-1)/2 + 1, C a multiple of 4.
 *   dd_maxpool3x3s2_nhwc_fwd : y and argmax (B,Ho,Wo,C bytes: position 0..8 of the maximum inside its window, first
 *                              maximum in scan order, NaN propagates -- the rule of ATen's kernel)
 *   dd_maxpool3x3s2_nhwc_bwd : grad_x (B,H,W,C) gathered from grad_y and argmax (every element written once; no atomics)
 * ------------------------------------------------------------------------------------------ */
int dd_maxpool3x3s2_nhwc_fwd(const float* x, int B, int H, int W, int C, float* y, unsigned char* argmax, void* stream);
int dd_maxpool3x3s2_nhwc_bwd(const float* grad_y, const unsigned char* argmax, int B, int H, int W, int C, float* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cross-covariance attention core of the Lite-Mono LGFI blocks (csrc/xca.cu): networks/depth_encoder.py:63-83 `XCA.forward`
 * between its qkv and proj linear layers.  qkv is the (B,N,3C) output of the qkv layer (q | k | v, each heads x d channels),
 * temperature is (heads); d = C / heads in {8, 16, 28}, C a multiple of 32 and <= 256.
 *   dd_xca_fwd : out (B,N,C) = (softmax_j((q^ k^T)[i][j] * temperature[head]) @ v) in token-major order, q^ / k^ = q / k
 *                L2-normalised over the N tokens (F.normalize, eps 1e-12); attn, scores (B,C,d) and rq, rk (B,C) = the
 *                attention rows, the normalised Gram rows and the reciprocal norms are kept for the backward pass
 *   dd_xca_bwd : grad_qkv (B,N,3C) and grad_temp_part (B,C): the temperature gradient of head h is the sum of
 *                grad_temp_part over the batch and the d channels of the head
 * Workspace: dd_xca_workspace_bytes(B, N, C, heads) (per-chunk partial Gram rows + coefficient matrices; deterministic).
 * ------------------------------------------------------------------------------------------ */
size_t dd_xca_workspace_bytes(int B, int N, int C, int heads);
int dd_xca_fwd(const float* qkv, const float* temperature, int B, int N, int C, int heads, float* out, float* attn, float* scores, float* rq,
               float* rk, void* workspace, size_t workspace_bytes, void* stream);
int dd_xca_bwd(const float* qkv, const float* temperature, const float* grad_out, const float* attn, const float* scores, const float* rq,
               const float* rk, int B, int N, int C, int heads, float* grad_qkv, float* grad_temp_part, void* workspace, size_t workspace_bytes,
               void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode BatchNorm2d + residual add + activation on channels_last activations (csrc/batchnorm_nhwc.cu): the
 * conv-bn-relu / conv-bn-(+identity)-relu pattern of the ResNet trunks (torchvision BasicBlock as used by
 * networks/resnet_encoder.py:16-20,:125-134) and the Lite-Mono stem's BNGELU (networks/depth_encoder.py:113-122).
 * x, residual, y, grad_* are (M = N*H*W, C) row-major (the memory of a channels_last tensor); C = 4 * a divisor of 256;
 * act: 0 none, 1 ReLU, 2 exact GELU; gamma / beta / residual may be NULL.
 *   dd_bn_act_nhwc_fwd : y = act((x - mean) * invstd * gamma + beta + residual) with batch statistics; save_mean / save_invstd
 *                        (C) for the backward pass; running statistics (may be NULL) updated as nn.BatchNorm2d does
 *   dd_bn_act_nhwc_bwd : grad_x, grad_residual (= grad_y * act'), grad_gamma, grad_beta (each may be NULL); y = the forward
 *                        output (needed for act = ReLU only: its sign is the mask)
 * Workspace: dd_bn_nhwc_workspace_bytes(C) (per-CTA partial sums; deterministic).
 * ------------------------------------------------------------------------------------------ */
size_t dd_bn_nhwc_workspace_bytes(int C);
int dd_bn_act_nhwc_fwd(const float* x, const float* residual, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       int act, float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* workspace,
                       size_t workspace_bytes, void* stream);
int dd_bn_act_nhwc_bwd(const float* x, const float* y, const float* grad_y, long long M, int C, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, int act, float* grad_x, float* grad_residual, float* grad_gamma,
                       float* grad_beta, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAMO_B200_H */
