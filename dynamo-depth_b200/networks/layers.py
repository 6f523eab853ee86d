"""Decoder building blocks with the reference's names (networks/layers.py:7-121), executed by the
hand-written sm_100a convolution kernels (dd_conv_*).  ReflectionPad2d + Conv2d + ELU (+ the x2
up-sampling and skip concatenation that precede them in the decoders) are one kernel launch.

The geometry layers the north-star lists under networks/layers.py live in `tools` in the reference
(tools.py:167-326); they are re-exported here so both import paths work.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from dd_b200.functional import conv2d_fused, pose_matrix, resize_bilinear
from tools import BackprojectDepth, Project3D, SSIM, compute_smooth_loss, disp_to_depth, depth_to_disp  # noqa: F401


def rot_from_axisangle(vec):
    """(B,1,3) axis-angle -> (B,4,4) rotation (Rodrigues; angle + 1e-7 in the axis normalisation,
    reference: networks/layers.py:43-82)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    rows = [x * xC + ca, xyC - zs, zxC + ys, zero,
            xyC + zs, y * yC + ca, yzC - xs, zero,
            zxC - ys, yzC + xs, z * zC + ca, zero,
            zero, zero, zero, one]
    return torch.cat(rows, 2).reshape(vec.shape[0], 4, 4)


def get_translation_matrix(translation_vector):
    """(B,1,3) or (B,3) -> (B,4,4) homogeneous translation (reference: networks/layers.py:27-40)."""
    t = translation_vector.contiguous().view(-1, 3, 1)
    B = t.shape[0]
    eye = torch.eye(4, device=t.device, dtype=t.dtype).expand(B, 4, 4)
    pad = torch.zeros(B, 4, 3, device=t.device, dtype=t.dtype)
    col = torch.cat([t, torch.zeros(B, 1, 1, device=t.device, dtype=t.dtype)], 1)
    return eye + torch.cat([pad, col], 2)


def transformation_from_parameters(axisangle, translation, invert=False):
    """network (axisangle, translation) -> 4x4; invert=True gives R^T @ Trans(-t) (reference: layers.py:7-24).
    CUDA tensors go through the fused dd_pose_matrix_* kernels (one launch instead of ~40 ATen ops)."""
    if axisangle.is_cuda:
        return pose_matrix(axisangle, translation, invert)
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


class Conv3x3(nn.Module):
    """pad (reflect | zero) + 3x3 convolution; parameters live in `self.conv` exactly as in the
    reference (state_dict key `conv.weight` / `conv.bias`)."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.use_refl = use_refl
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x, skip=None, up="none", act="none"):
        return conv2d_fused(x, self.conv.weight, self.conv.bias, x1=skip, ksize=3,
                            pad="reflect" if self.use_refl else "zero", act=act, up=up)


class ConvBlock(nn.Module):
    """Conv3x3 + ELU (reference: networks/layers.py:85-97).  `skip` / `up` let the decoders fuse the
    preceding `upsample` + `torch.cat` into the same launch."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)   # kept for module-tree parity; applied inside the kernel

    def forward(self, x, skip=None, up="none"):
        return self.conv(x, skip=skip, up=up, act="elu")


def upsample(x, scale_factor=2, mode="nearest"):
    """Stand-alone x2 up-sampling (reference: networks/layers.py:118-121).  The decoders never call it
    (fused into the next convolution); bilinear goes through the resize kernel."""
    if mode == "bilinear" and x.is_cuda:
        return resize_bilinear(x, (x.shape[-2] * scale_factor, x.shape[-1] * scale_factor))
    return F.interpolate(x, scale_factor=scale_factor, mode=mode)
