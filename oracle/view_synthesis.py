"""ORACLE (test infrastructure, never on the product path) -- CPU restatement of the
reference's differentiable view-synthesis loss.

Each function restates, with elementary torch tensor ops on whatever dtype it is given
(fp32 for parity, fp64 for tolerance arbitration), the arithmetic of one reference
function and cites it (paths relative to /root/reference, commit 227a5d9).  Gradients
come from torch autograd over these elementary ops.  The restatement is pinned against
golden vectors produced by executing the reference itself (oracle/gen_golden.py ->
tests/golden/*.npz, checked by tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package.
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a1  utils.interp (utils.py:98-101)  ==  F.interpolate(mode='bilinear', align_corners=False)
#     semantics restated from ATen UpSample.h: scale = in/out, src = scale*(dst+0.5)-0.5
#     clamped to >= 0, second tap = min(i0+1, n-1); same size is a plain copy.
# --------------------------------------------------------------------------------------


def _axis_taps(n_in, n_out, dtype, device):
    if n_in == n_out:
        idx = torch.arange(n_out, device=device)
        return idx, idx, torch.zeros(n_out, dtype=dtype, device=device)
    # ATen computes the source index in the tensor's own precision ("accscalar" = float for fp32)
    scale = torch.tensor(n_in / n_out, dtype=dtype, device=device)
    dst = torch.arange(n_out, dtype=dtype, device=device)
    src = (scale * (dst + 0.5) - 0.5).clamp_min(0)
    i0 = src.floor().to(torch.long)
    i1 = torch.clamp(i0 + 1, max=n_in - 1)
    lam = src - i0.to(dtype)
    return i0, i1, lam


def bilinear_resize(x, size):
    """(B,C,h,w) -> (B,C,*size); utils.py:98-101."""
    h_in, w_in = x.shape[-2:]
    h_out, w_out = size
    if (h_in, w_in) == (h_out, w_out):
        return x.clone()
    y0, y1, ly = _axis_taps(h_in, h_out, x.dtype, x.device)
    x0, x1, lx = _axis_taps(w_in, w_out, x.dtype, x.device)
    ly = ly.view(-1, 1)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    rows = top * (1 - ly) + bot * ly  # (B,C,h_out,w_in)
    left = rows[..., x0]
    right = rows[..., x1]
    return left * (1 - lx) + right * lx


# --------------------------------------------------------------------------------------
# a2  disp_to_depth / depth_to_disp (tools.py:291-308)
# --------------------------------------------------------------------------------------


def disp_to_depth(disp, min_depth, max_depth):
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def depth_to_disp(depth, min_depth, max_depth):
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    return (1 / depth - min_disp) / (max_disp - min_disp)


# --------------------------------------------------------------------------------------
# a3  BackprojectDepth (tools.py:167-197): P = (depth * inv_K[:3,:3] @ (u,v,1), 1)
# --------------------------------------------------------------------------------------


def pixel_grid(height, width, dtype, device):
    """(3, H*W) rows (u, v, 1), row-major index v*W+u (tools.py:177-189)."""
    v, u = torch.meshgrid(
        torch.arange(height, dtype=dtype, device=device),
        torch.arange(width, dtype=dtype, device=device),
        indexing="ij",
    )
    return torch.stack([u.reshape(-1), v.reshape(-1), torch.ones(height * width, dtype=dtype, device=device)], 0)


def backproject(depth, inv_K):
    """depth (B,1,H,W), inv_K (B,4,4) -> (B,4,H*W)."""
    B, _, H, W = depth.shape
    pix = pixel_grid(H, W, depth.dtype, depth.device)
    rays = torch.matmul(inv_K[:, :3, :3], pix.unsqueeze(0).expand(B, -1, -1))
    pts = depth.reshape(B, 1, -1) * rays
    return torch.cat([pts, torch.ones(B, 1, H * W, dtype=depth.dtype, device=depth.device)], 1)


# --------------------------------------------------------------------------------------
# a4  Project3D (tools.py:200-224)
# --------------------------------------------------------------------------------------


def project(points, K, T, height, width, eps=1e-7):
    """points (B,4,P) -> (normalised grid (B,H,W,2), ego_motion (B,3,P)).  T may be None."""
    X = torch.matmul(T, points) if T is not None else points
    c = torch.matmul(K[:, :3, :], X)
    pix = c[:, :2, :] / (c[:, 2:3, :] + eps)
    B = points.shape[0]
    pix = pix.reshape(B, 2, height, width).permute(0, 2, 3, 1)
    gx = pix[..., 0] / (width - 1)
    gy = pix[..., 1] / (height - 1)
    grid = (torch.stack([gx, gy], -1) - 0.5) * 2
    return grid, X[:, :3] - points[:, :3]


# --------------------------------------------------------------------------------------
# a6  F.grid_sample(bilinear, padding_mode='border', align_corners=True)  (Trainer.py:281)
#     restated from ATen GridSampler.cuh: un-normalise ((g+1)/2)*(size-1); clip to
#     [0,size-1] with ZERO coordinate gradient when the un-clipped value is <=0 or >=size-1;
#     corners floor/ +1; a corner outside the image contributes nothing.
# --------------------------------------------------------------------------------------


def _clip_with_aten_grad(coord, size):
    hi = float(size - 1)
    clipped = coord.clamp(0.0, hi)
    live = ((coord > 0) & (coord < hi)).to(coord.dtype)
    # value == clipped; d/dcoord == live (ATen clip_coordinates_set_grad)
    return clipped.detach() + (coord - coord.detach()) * live


def grid_sample_border(img, grid):
    """img (B,C,H,W), grid (B,Ho,Wo,2) in [-1,1] -> (B,C,Ho,Wo)."""
    B, C, H, W = img.shape
    ix = _clip_with_aten_grad((grid[..., 0] + 1) / 2 * (W - 1), W)
    iy = _clip_with_aten_grad((grid[..., 1] + 1) / 2 * (H - 1), H)
    x0f = ix.detach().floor()
    y0f = iy.detach().floor()
    tx = ix - x0f
    ty = iy - y0f
    x0 = x0f.to(torch.long)
    y0 = y0f.to(torch.long)
    x1 = x0 + 1
    y1 = y0 + 1
    x1_ok = (x1 <= W - 1).to(img.dtype)
    y1_ok = (y1 <= H - 1).to(img.dtype)
    x1c = x1.clamp(max=W - 1)
    y1c = y1.clamp(max=H - 1)

    flat = img.reshape(B, C, H * W)

    def tap(yy, xx):
        idx = (yy * W + xx).reshape(B, 1, -1).expand(-1, C, -1)
        return torch.gather(flat, 2, idx).reshape(B, C, *grid.shape[1:3])

    nw = tap(y0, x0)
    ne = tap(y0, x1c) * x1_ok.unsqueeze(1)
    sw = tap(y1c, x0) * y1_ok.unsqueeze(1)
    se = tap(y1c, x1c) * (x1_ok * y1_ok).unsqueeze(1)
    tx = tx.unsqueeze(1)
    ty = ty.unsqueeze(1)
    return nw * (1 - tx) * (1 - ty) + ne * tx * (1 - ty) + sw * (1 - tx) * ty + se * tx * ty


# --------------------------------------------------------------------------------------
# a7  SSIM (tools.py:227-257): reflect pad 1, 3x3 box means, C1=0.01^2, C2=0.03^2
# --------------------------------------------------------------------------------------


def reflect_pad1(x):
    x = torch.cat([x[..., 1:2, :], x, x[..., -2:-1, :]], -2)
    return torch.cat([x[..., :, 1:2], x, x[..., :, -2:-1]], -1)


def box3_mean(xp):
    """3x3 stride-1 'valid' mean of a padded map (AvgPool2d(3,1), tools.py:232-236)."""
    H, W = xp.shape[-2] - 2, xp.shape[-1] - 2
    acc = None
    for dy in range(3):
        for dx in range(3):
            t = xp[..., dy : dy + H, dx : dx + W]
            acc = t if acc is None else acc + t
    return acc / 9


def ssim(x, y):
    C1, C2 = 0.01**2, 0.03**2
    x = reflect_pad1(x)
    y = reflect_pad1(y)
    mu_x = box3_mean(x)
    mu_y = box3_mean(y)
    sigma_x = box3_mean(x * x) - mu_x * mu_x
    sigma_y = box3_mean(y * y) - mu_y * mu_y
    sigma_xy = box3_mean(x * y) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x * mu_x + mu_y * mu_y + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


# a8  compute_reprojection_loss (Trainer.py:413-423)
def reprojection_loss(pred, target, ssim_weight=0.85):
    l1 = (target - pred).abs().mean(1, True)
    s = ssim(pred, target).mean(1, True)
    return ssim_weight * s + (1 - ssim_weight) * l1


# a10 compute_smooth_loss (tools.py:311-326)
def smooth_loss(inp, img=None):
    gx = (inp[:, :, :, :-1] - inp[:, :, :, 1:]).abs()
    gy = (inp[:, :, :-1, :] - inp[:, :, 1:, :]).abs()
    if img is not None:
        wx = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
        wy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
        gx = gx * torch.exp(-wx)
        gy = gy * torch.exp(-wy)
    return gx.mean() + gy.mean()


# --------------------------------------------------------------------------------------
# Configuration carrier (the subset of options.py / Trainer state the loss path reads)
# --------------------------------------------------------------------------------------

LOSS_TERMS = ["p_photo", "d_smooth", "d_ground", "c_smooth", "c_consistency", "m_sparsity", "m_smooth"]


class LossConfig:
    """Defaults are options.py:78-118,182-189; phase flags follow Trainer.py:466-490."""

    def __init__(self, height, width, scales, frame_ids=(0, -1, 1), phase="disp_init", **kw):
        self.height, self.width = height, width
        self.scales = list(scales)
        self.frame_ids = list(frame_ids)
        self.min_depth, self.max_depth = 0.1, 100.0
        self.ssim_weight = 0.85
        self.mask_disp_thrd = 0.03
        self.g = dict(p_photo=1.0, d_smooth=1e-3, d_ground=0.1, c_smooth=1e-3, c_consistency=5.0,
                      m_sparsity=0.04, m_smooth=0.1)
        self.weight_ramp = ["g_c_smooth", "g_c_consistency", "g_m_sparsity", "g_m_smooth"]
        self.ramp_red = 3.0
        self.gp_prior, self.gp_tol, self.gp_max_it, self.gp_np_per_it = 0.4, 0.005, 100, 5
        self.set_phase(phase)
        for k, v in kw.items():
            if k.startswith("g_"):
                self.g[k[2:]] = v
            else:
                setattr(self, k, v)

    def set_phase(self, phase):
        table = {
            "disp_init": (False, False, ["Depth", "Pose"]),
            "motion_init": (True, False, ["CmpFlow"]),
            "mask_init": (True, True, ["Pose", "CmpFlow", "MotMask"]),
            "fine_tune": (True, True, ["Depth", "Pose", "CmpFlow", "MotMask"]),
        }
        self.phase = phase
        self.bool_CmpFlow, self.bool_MotMask, self.network_names = table[phase]
        self.automask = phase == "disp_init"  # Trainer.py:117

    def coefficients(self, step, steps_per_epoch):
        """Trainer.py:302-310 (ramp = clip(ramp_red*step/steps_per_epoch, 0, 1))."""
        out = {}
        for term in LOSS_TERMS:
            val = self.g[term]
            if "g_" + term in self.weight_ramp:
                val = val * float(np.clip(self.ramp_red * step / steps_per_epoch, 0.0, 1.0))
            out[term] = val
        return out


# --------------------------------------------------------------------------------------
# a5/a6  Trainer.generate_images_pred (Trainer.py:215-287)
# --------------------------------------------------------------------------------------


def generate_images_pred(cfg, inputs, outputs):
    H, W = cfg.height, cfg.width
    for s in cfg.scales:
        disp_s = outputs[("disp", 0, s)]
        B = disp_s.shape[0]
        h, w = disp_s.shape[-2:]
        disp = bilinear_resize(disp_s, (H, W))
        scaled, depth = disp_to_depth(disp, cfg.min_depth, cfg.max_depth)
        outputs[("depth", 0, s)] = depth
        outputs[("disp_scaled", 0, s)] = scaled
        K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
        for f in cfg.frame_ids[1:]:
            T = outputs[("cam_T_cam", 0, f)]
            cam = backproject(depth, inv_K)
            outputs[("cam_points", 0, s)] = cam
            if cfg.bool_MotMask:
                mask_r = bilinear_resize(outputs[("motion_mask", f, s)], (H, W))
            else:
                outputs[("motion_mask", f, s)] = torch.ones(B, 1, h, w, dtype=disp_s.dtype)
                mask_r = torch.ones(B, 1, H, W, dtype=disp_s.dtype)
            outputs[("motion_mask_r", f, s)] = mask_r

            if cfg.bool_CmpFlow:
                sample_ego, ego_flow = project(cam, K, T, H, W)
                ts = inputs[("ts", f)].reshape(B, 1, 1)
                complete = bilinear_resize(outputs[("complete_flow", f, s)], (H, W)).reshape(B, 3, -1) * ts
                residual = complete - ego_flow
                independ = residual * mask_r.reshape(B, 1, -1)
                outputs[("sample_ego", f, s)] = sample_ego.detach()
                tmp = cam.detach().clone()
                tmp = torch.cat([tmp[:, :3] + complete, tmp[:, 3:]], 1)
                sample_complete, _ = project(tmp, K, None, H, W)
                outputs[("sample_complete", f, s)] = sample_complete.detach()
                if cfg.bool_MotMask:
                    cam2 = backproject(depth, inv_K)
                    cam2 = torch.cat([cam2[:, :3] + independ, cam2[:, 3:]], 1)
                    sample, _ = project(cam2, K, T, H, W)
                else:
                    cam2 = torch.cat([cam[:, :3] + complete, cam[:, 3:]], 1)
                    sample, _ = project(cam2, K, None, H, W)
            else:
                sample, ego_flow = project(cam, K, T, H, W)
                residual = torch.zeros_like(ego_flow)
                independ = torch.zeros_like(ego_flow)

            outputs[("sample", f, s)] = sample
            outputs[("color", f, s)] = grid_sample_border(inputs[("color", f, 0)], sample)
            outputs[("ego_flow", f, s)] = ego_flow
            outputs[("independ_flow", f, s)] = independ.reshape(B, 3, H, W)
            outputs[("residual_flow", f, s)] = bilinear_resize(residual.reshape(B, 3, H, W), (h, w))
            if cfg.automask:
                outputs[("color_identity", f, s)] = inputs[("color", f, 0)]


# --------------------------------------------------------------------------------------
# a9/a11/a12  Trainer.compute_losses (Trainer.py:289-411)
# --------------------------------------------------------------------------------------


def softplus_mean(x):
    """BCEWithLogitsLoss(x, 0) == mean(softplus(x)) (Trainer.py:66,399)."""
    return (x.clamp_min(0) + torch.log1p(torch.exp(-x.abs()))).mean()


def compute_losses(cfg, inputs, outputs, step, steps_per_epoch, noise=None, ground_fn=None):
    """noise: dict scale -> (B, F, H, W) standard-normal tie-break tensor replacing torch.randn at
    Trainer.py:339 (the reference draws it inside; parity needs it injected).
    ground_fn(inputs, outputs, scale) -> disp_diff implements Trainer.process_ground when supplied."""
    move_Depth = "Depth" in cfg.network_names
    move_CmpFlow = "CmpFlow" in cfg.network_names
    move_MotMask = "MotMask" in cfg.network_names
    coef = cfg.coefficients(step, steps_per_epoch)
    losses = {"loss": 0}
    for t in LOSS_TERMS + cfg.scales:
        losses[f"loss_term/{t}"] = 0
    for t in LOSS_TERMS:
        losses[f"loss_coef/{t}"] = coef[t]
    sources = cfg.frame_ids[1:]
    nf = len(sources)
    target = inputs[("color", 0, 0)]
    for s in cfg.scales:
        ps = {t: 0 for t in LOSS_TERMS}
        color = inputs[("color", 0, s)]
        reproj = torch.cat([reprojection_loss(outputs[("color", f, s)], target, cfg.ssim_weight) for f in sources], 1)
        if cfg.automask:
            ident = torch.cat([reprojection_loss(inputs[("color", f, 0)], target, cfg.ssim_weight) for f in sources], 1)
            if noise is not None:
                ident = ident + noise[s].to(ident.dtype) * 0.00001
            combined = torch.cat([ident, reproj], 1)
        else:
            combined = reproj
        if combined.shape[1] == 1:
            to_opt = combined
        else:
            to_opt, idxs = torch.min(combined, dim=1)
        if cfg.automask:
            outputs[f"identity_selection/{s}"] = (idxs > nf - 1).to(target.dtype)
        ps["p_photo"] = to_opt.mean()

        if move_Depth:
            if coef["d_smooth"] > 0:
                disp = outputs[("disp", 0, s)]
                norm_disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
                ps["d_smooth"] = smooth_loss(norm_disp, color) / (2**s)
            if coef["d_ground"] > 0 and cfg.bool_MotMask and ground_fn is not None:
                disp_diff = ground_fn(inputs, outputs, s)
                disp_diff = torch.where(disp_diff > 0, torch.zeros_like(disp_diff), disp_diff)
                ps["d_ground"] = -1 * disp_diff.mean() / (2**s)

        for f in sources:
            disp = outputs[("disp", 0, s)]
            mask = outputs[("motion_mask", f, s)]
            h, w = mask.shape[-2:]
            if move_CmpFlow and cfg.bool_CmpFlow:
                if coef["c_smooth"] > 0:
                    ps["c_smooth"] = ps["c_smooth"] + smooth_loss(outputs[("complete_flow", f, s)], color) / (2**s) / nf
                if cfg.bool_MotMask and coef["c_consistency"] > 0:
                    valid = (disp > cfg.mask_disp_thrd).detach().to(disp.dtype)
                    res = outputs[("residual_flow", f, s)]
                    ps["c_consistency"] = ps["c_consistency"] + (valid * (1 - mask.detach()) * res.abs()).mean() / (2**s) / nf
            if move_MotMask and cfg.bool_MotMask:
                if coef["m_sparsity"] > 0:
                    se = bilinear_resize(outputs[("sample_ego", f, s)].permute(0, 3, 1, 2), (h, w))
                    sc = bilinear_resize(outputs[("sample_complete", f, s)].permute(0, 3, 1, 2), (h, w))
                    mag = ((se - sc) ** 2).sum(1)
                    static = (mag < mag.mean()).unsqueeze(1)
                    if bool(torch.all(static.sum((1, 2, 3)) > 0)):
                        ps["m_sparsity"] = ps["m_sparsity"] + softplus_mean(outputs[("motion_prob", f, s)][static]) / (2**s) / nf
                if coef["m_smooth"] > 0:
                    ps["m_smooth"] = ps["m_smooth"] + smooth_loss(mask, color) / (2**s) / nf

        for t in LOSS_TERMS:
            losses[f"loss_term/{s}"] = losses[f"loss_term/{s}"] + ps[t] * coef[t]
            losses[f"loss_term/{t}"] = losses[f"loss_term/{t}"] + ps[t]
        losses["loss"] = losses["loss"] + losses[f"loss_term/{s}"] / len(cfg.scales)
    return losses
