"""ORACLE (test infrastructure) -- CPU restatement of the ground-plane prior of phase fine_tune:
tools.GroundPlane (tools.py:76-164) and Trainer.process_ground / get_ground_depth (Trainer.py:425-461).

`rand_index_fn(n, k)` replaces the reference's host RNG draw np.random.choice(np.arange(N), T, replace=True)
(tools.py:126) so the hypotheses can be reproduced.
"""
import numpy as np
import torch

from . import view_synthesis as vs


def _design(pts):
    """[x, z, 1] and y for vertical_axis = 1 (tools.py:155-164)."""
    return torch.cat([pts[..., 0:1], pts[..., 2:3], torch.ones_like(pts[..., 1:2])], -1), pts[..., 1:2]


def estimate_ground_plane(ground, num_points_per_it, max_it, tol, rand_index_fn):
    """ground (B,N,3) -> best plane parameters (B,3,1) (tools.py:113-139)."""
    B, N, _ = ground.shape
    k = num_points_per_it * max_it
    picks = torch.stack([ground[b][rand_index_fn(N, k)] for b in range(B)])
    A, rhs = _design(picks.reshape(-1, num_points_per_it, 3))
    At = A.transpose(2, 1)
    ws = (torch.inverse(At @ A + 1e-6) @ At @ rhs).reshape(-1, 3, 1)   # (B*max_it,3,1), image-major
    # hypothesis k is scored on the points of image k % B: the reference tiles with .repeat(max_it,1,1) (tools.py:131)
    A_all, y_all = _design(ground.repeat(max_it, 1, 1))                  # (B*max_it, N, 3), (B*max_it, N, 1)
    counts = ((A_all @ ws - y_all).abs() < tol).float().mean((1, 2))
    best = counts.reshape(B, max_it).argmax(1)
    return ws.reshape(B, max_it, 3, 1)[torch.arange(B), best]


def ground_plane(points, g_prior, num_points_per_it, max_it, tol, rand_index_fn):
    """points (B,3,H,W) -> (vertical distance to the plane (B,1,H,W), plane parameters (B,3,1)), detached."""
    B, _, H, W = points.shape
    ground = points[:, :, -int(g_prior * H):, :].reshape(B, 3, -1).permute(0, 2, 1)
    param = estimate_ground_plane(ground, num_points_per_it, max_it, tol, rand_index_fn)
    A, y = _design(points.reshape(B, 3, H * W).permute(0, 2, 1))
    dist = (A @ param - y).permute(0, 2, 1).reshape(B, 1, H, W)
    return dist.detach(), param.detach()


def make_ground_fn(cfg, rand_index_fn):
    """ground_fn(inputs, outputs, scale) -> disp_diff for oracle.view_synthesis.compute_losses (Trainer.py:425-461)."""

    def ground_fn(inputs, outputs, s):
        disp = outputs[("disp", 0, s)]
        _, depth = vs.disp_to_depth(disp, cfg.min_depth, cfg.max_depth)
        inv_K = inputs[("inv_K", s)]
        B, _, h, w = disp.shape
        cam = vs.backproject(depth, inv_K)
        _, param = ground_plane(cam[:, :3].reshape(B, 3, h, w), cfg.gp_prior, cfg.gp_np_per_it, cfg.gp_max_it, cfg.gp_tol,
                                rand_index_fn)
        param = param.clone()
        param[:, 2] += cfg.gp_tol
        rays = torch.matmul(inv_K[:, :3, :3], vs.pixel_grid(h, w, disp.dtype, disp.device).unsqueeze(0).expand(B, -1, -1))
        w1, w2, w3 = param[:, 0:1], param[:, 1:2], param[:, 2:3]
        vx, vy, vz = rays[:, 0:1], rays[:, 1:2], rays[:, 2:3]
        gd = (w3 / (vy - vx * w1 - vz * w2)).reshape(B, 1, h, w)
        gd = torch.where((gd < 0) | (gd > cfg.max_depth), torch.full_like(gd, cfg.max_depth), gd)
        ground_disp = vs.depth_to_disp(gd, cfg.min_depth, cfg.max_depth)
        diff = disp - ground_disp
        return torch.where(gd == cfg.max_depth, torch.zeros_like(diff), diff)

    return ground_fn


class SeededIndices:
    """Deterministic stand-in for np.random.choice(np.arange(N), T, replace=True): call i returns
    RandomState(seed + i).randint(0, n, k)."""

    def __init__(self, seed):
        self.seed, self.calls = seed, 0

    def __call__(self, n, k):
        out = np.random.RandomState(self.seed + self.calls).randint(0, n, k)
        self.calls += 1
        return out
