"""Geometry / photometric layers with the reference's names and signatures (tools.py:167-326).

The training step never calls these one by one: Trainer.generate_images_pred / compute_losses go
through the fused kernels (dd_b200.functional.view_synthesis_sums).  These classes are the module
surface eval scripts and user code rely on (`trainer.backproject_depth[s](depth, inv_K)`, ...), each
backed by its own sm_100a kernel.  CPU tensors are refused (no fallback) except in the host-side
monitoring helpers (DepthMetrics, GroundPlane) which are not part of the hot path.
"""
import numpy as np
import torch
import torch.nn as nn

from dd_b200 import functional as _F


class BackprojectDepth(nn.Module):
    """depth image -> homogeneous camera points (B,4,H*W) (reference: tools.py:167-197)."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        assert depth.shape[-2:] == (self.height, self.width), (depth.shape, self.height, self.width)
        return _F.backproject(depth, inv_K)


class Project3D(nn.Module):
    """camera points -> normalised sampling grid + ego-motion field (reference: tools.py:200-224)."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        if eps != 1e-7:
            raise NotImplementedError("Project3D: eps is fixed to 1e-7 in the kernels (tools.py:203)")
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        return _F.project3d(points, K, T, self.height, self.width)


class SSIM(nn.Module):
    """clamp((1 - SSIM(x, y)) / 2, 0, 1) with 3x3 reflect-padded windows (reference: tools.py:227-257)."""

    def forward(self, x, y):
        return _F.ssim(x, y)


def disp_to_depth(disp, min_depth, max_depth):
    """sigmoid output -> (scaled disparity, depth) (reference: tools.py:291-298); two scalar-affine ops,
    fused into dd_warp_photo_* on the training path."""
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def depth_to_disp(depth, min_depth, max_depth):
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    return (1 / depth - min_disp) / (max_disp - min_disp)


def compute_smooth_loss(inp, img=None):
    """edge-aware first-order smoothness, mean_x + mean_y (reference: tools.py:311-326)."""
    sums = _F.smooth_sums([inp], [img], [False])
    return _F.smooth_means(sums, [tuple(inp.shape)])[0]


# ---------------------------------------------------------------------------------------------------
# host-side monitoring helpers (out of the kernel scope; plain torch, any device)
# ---------------------------------------------------------------------------------------------------


def torch_and(*args):
    out = args[0]
    for a in args[1:]:
        assert out.size() == a.size(), "Sizes must match"
        out = torch.logical_and(out, a)
    return out


def compute_errors(gt, pred):
    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = [(ratio < 1.25**k).float().mean() for k in (1, 2, 3)]
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = ((gt - pred).abs() / gt).mean()
    sq_rel = (((gt - pred) ** 2) / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


class DepthMetrics(nn.Module):
    """LiDAR depth metrics with median scaling on validation batches (reference: tools.py:6-73)."""

    names = ["de:abs_rel", "de:sq_rel", "de:rms", "de:log_rms", "da:a1", "da:a2", "da:a3"]

    def __init__(self, img_bound, min_depth, max_depth):
        super().__init__()
        self.depth_metric_names = list(self.names)
        self.img_bound, self.min_depth, self.max_depth = img_bound, min_depth, max_depth

    def forward(self, inputs, outputs, mask=None):
        disp_pred = outputs[("disp_scaled", 0, 0)]
        totals = {k: 0 for k in self.depth_metric_names}
        for bi in range(disp_pred.shape[0]):
            pts, valid = inputs["depth_gt"][bi], inputs["depth_valid"][bi]
            gh, gw = int(inputs["gt_dim"][bi][0]), int(inputs["gt_dim"][bi][1])
            top, bot = int(self.img_bound[0] * gh), int(self.img_bound[1] * gh)
            lft, rgt = int(self.img_bound[2] * gw), int(self.img_bound[3] * gw)
            valid = torch_and(valid, pts[:, 0] >= top, pts[:, 0] < bot, pts[:, 1] >= lft, pts[:, 1] < rgt,
                              pts[:, 2] > self.min_depth, pts[:, 2] < self.max_depth)
            rows, cols = pts[:, 0][valid].long(), pts[:, 1][valid].long()
            full = nn.functional.interpolate(disp_pred[bi][None], (gh, gw), mode="bilinear", align_corners=False).squeeze()
            d_gt, d_pd = pts[:, 2][valid], (1 / full)[rows, cols]
            d_pd = torch.clamp(d_pd * (torch.median(d_gt) / torch.median(d_pd)), self.min_depth, self.max_depth)
            for k, v in zip(self.depth_metric_names, compute_errors(d_gt, d_pd)):
                totals[k] = totals[k] + v
        return {k: v / disp_pred.shape[0] for k, v in totals.items()}


class GroundPlane(nn.Module):
    """RANSAC ground-plane fit used by the d_ground prior in phase fine_tune (reference: tools.py:76-164).
    Hypothesis sampling keeps the reference's host numpy RNG (injectable through `rand_index_fn` for parity
    tests); the hypothesis scoring -- B*max_it planes against 0.4*H*W points, a 1.9 GB intermediate at bs32 in
    the reference -- is one dd_ground_score launch that reads every point once per 20 hypotheses."""

    def __init__(self, num_points_per_it=5, max_it=25, tol=0.1, g_prior=0.5, vertical_axis=1, rand_index_fn=None):
        super().__init__()
        self.num_points_per_it, self.max_it, self.tol, self.g_prior = num_points_per_it, max_it, tol, g_prior
        self.vertical_axis = vertical_axis
        self.rand_index_fn = rand_index_fn or (lambda n, k: np.random.choice(np.arange(n), k, replace=True))

    def _design(self, pts):
        va = self.vertical_axis
        rhs = pts[..., va:va + 1]
        cols = [pts[..., i:i + 1] for i in range(3) if i != va] + [torch.ones_like(rhs)]
        return torch.cat(cols, -1), rhs

    def dist_from_plane(self, pts, param):
        A, rhs = self._design(pts)
        return A @ param - rhs

    _PIN_SLOTS = 8

    def _stage_indices(self, idx_np, device):
        """Host RNG draw -> device without a host synchronisation: a small ring of pinned buffers and non-blocking copies
        (a pageable cudaMemcpy blocks the host until the GPU reaches it, i.e. once per scale in the middle of the step);
        a slot is reused only after the copy issued from it has completed."""
        ring = self.__dict__.setdefault("_pin_ring", {})
        key = (tuple(idx_np.shape), str(device))
        slots = ring.get(key)
        if slots is None:
            slots = ring[key] = {"next": 0, "bufs": [torch.empty(idx_np.shape, dtype=torch.int64, pin_memory=True) for _ in range(self._PIN_SLOTS)],
                                 "events": [None] * self._PIN_SLOTS}
        i = slots["next"]
        slots["next"] = (i + 1) % self._PIN_SLOTS
        if slots["events"][i] is not None:
            slots["events"][i].synchronize()
        slots["bufs"][i].numpy()[...] = idx_np
        out = slots["bufs"][i].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        slots["events"][i] = ev
        return out

    def estimate_ground_plane_fused(self, points, row0):
        """points (B,3,H,W) on the GPU; same result as estimate_ground_plane on points[:, :, row0:, :]."""
        B, _, H, W = points.shape
        N = (H - row0) * W
        k = self.num_points_per_it * self.max_it
        flat = points[:, :, row0:, :].reshape(B, 3, N)
        idx = self._stage_indices(np.stack([np.asarray(self.rand_index_fn(N, k)) for _ in range(B)]), points.device)
        picks = torch.gather(flat, 2, idx.unsqueeze(1).expand(B, 3, k)).permute(0, 2, 1)        # (B, k, 3)
        A, rhs = self._design(picks.reshape(-1, self.num_points_per_it, 3))
        At = A.transpose(2, 1)
        # inv_ex: the same LU-based inverse as torch.inverse (tools.py:124) without its device -> host error check, which
        # would stall the host in the middle of every step
        ws = (torch.linalg.inv_ex(At @ A + 1e-6).inverse @ At @ rhs).reshape(-1, 3)            # (B*max_it, 3)
        counts = _F.ground_score(points, ws, row0, self.tol).reshape(B, self.max_it)
        best = counts.argmax(1)
        return ws.reshape(B, self.max_it, 3, 1)[torch.arange(B, device=points.device), best]

    def estimate_ground_plane(self, pts):
        B, N, _ = pts.shape
        k = self.num_points_per_it * self.max_it
        picks = torch.stack([pts[b][self.rand_index_fn(N, k)] for b in range(B)])          # (B, k, 3)
        A, rhs = self._design(picks.reshape(-1, self.num_points_per_it, 3))
        At = A.transpose(2, 1)
        ws = (torch.inverse(At @ A + 1e-6) @ At @ rhs).reshape(-1, 3, 1)                  # (B*max_it, 3, 1)
        # NB: the reference tiles the points with .repeat(max_it,1,1) (image-major order) against
        # hypotheses in (image, iteration) order (tools.py:131-133); reproduced as is for parity.
        tiled = pts.repeat(self.max_it, 1, 1)
        dist = self.dist_from_plane(tiled, ws).abs().reshape(B, self.max_it, N)
        best = (dist < self.tol).float().mean(2).argmax(1)
        return ws.reshape(B, self.max_it, 3, 1)[np.arange(B), best]

    def forward(self, points):
        B, _, H, W = points.shape
        if points.is_cuda and self.vertical_axis == 1:
            param = self.estimate_ground_plane_fused(points.detach().contiguous(), H - int(self.g_prior * H))
        else:
            ground = points[:, :, -int(self.g_prior * H):, :].reshape(B, 3, -1).permute(0, 2, 1)
            param = self.estimate_ground_plane(ground)
        dist = self.dist_from_plane(points.reshape(B, 3, H * W).permute(0, 2, 1), param).permute(0, 2, 1).reshape(B, 1, H, W)
        return dist.detach(), param.detach()
