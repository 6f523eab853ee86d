"""Host helpers keeping the names Trainer / eval code import from the reference's utils.py.

Only `interp` (utils.py:98-101 in the reference) is on the training hot path; there it is folded
into the fused kernels and this function is the stand-alone, module-surface version.  The
visualisation / video helpers of the reference (imageio, matplotlib) are out of scope.
"""
import os
import os.path as osp

import torch


def readlines(filename):
    with open(filename, "r") as fh:
        return fh.read().splitlines()


def write_to_file(data_list, fname, bool_newline=True):
    with open(fname, "w") as fh:
        fh.writelines([d + "\n" for d in data_list] if bool_newline else data_list)


def join_dir(*parts):
    """osp.join + makedirs (tolerates concurrent creation by other ranks)."""
    path = osp.join(*parts)
    os.makedirs(path, exist_ok=True)
    return path


def sec_to_hm_str(t):
    t = int(t)
    return f"{t // 3600:02d}h{(t % 3600) // 60:02d}m{t % 60:02d}s"


def interp(x, shape, mode="bilinear", align_corners=False):
    """(B,C,H,W) -> (B,C,*shape), F.interpolate(bilinear, align_corners=False) semantics.

    CUDA tensors go through the hand-written resize kernel (dd_resize_bilinear_*); anything else the
    reference never asks for on this path is refused rather than silently routed to PyTorch."""
    if mode != "bilinear" or align_corners:
        raise NotImplementedError("interp: only mode='bilinear', align_corners=False is part of the hot path")
    if x.is_cuda:
        from dd_b200.functional import resize_bilinear

        return resize_bilinear(x, shape)
    # CPU tensors only occur in data preparation / evaluation code, never in the training step
    return torch.nn.functional.interpolate(x, shape, mode=mode, align_corners=align_corners)
