"""Seeded synthetic frame triplets with the reference's `inputs` dict layout (datasets/base_dataset.py:
99-204): low-pass random textures, source frames = target translated by (+-3, +-1)
pixels plus noise, camera intrinsics of the KITTI / Waymo / nuScenes rigs scaled to the image size.
Used by bench.py, __graft_entry__.smoke() and Trainer.train() when no dataset package is supplied."""
import numpy as np
import torch
import torch.nn.functional as F

_NORM_K = {"kitti": (0.58, 1.92, 0.5, 0.5), "waymo": (1.06485, 1.59727, 0.49457, 0.49472),
           "nuscenes": (0.79151, 1.40713, 0.51017, 0.54612)}


def camera(kind, height, width, batch):
    fx, fy, cx, cy = _NORM_K[kind]
    K = np.array([[fx * width, 0, cx * width, 0], [0, fy * height, cy * height, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    inv_K = np.linalg.pinv(K).astype(np.float32)
    rep = lambda a: torch.from_numpy(a).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    return rep(K), rep(inv_K)


def _texture(gen, shape):
    x = torch.rand(shape, generator=gen)
    x = F.avg_pool2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), 7, 1)
    x = x - x.amin(dim=(1, 2, 3), keepdim=True)
    return x / x.amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)


def _shift(img, dx, dy):
    H, W = img.shape[-2:]
    ys = (torch.arange(H) + dy).clamp(0, H - 1)
    xs = (torch.arange(W) + dx).clamp(0, W - 1)
    return img[:, :, ys][:, :, :, xs]


def make_batch(opt, seed, batch=None, kind=None):
    """One `inputs` dict of CPU tensors (what a DataLoader worker would hand over)."""
    B = batch or opt.batch_size
    H, W = opt.height, opt.width
    kind = kind or opt.dataset
    gen = torch.Generator().manual_seed(seed)
    inputs = {}
    tgt = _texture(gen, (B, 3, H, W))
    inputs[("color", 0, 0)] = tgt
    for f in opt.frame_ids[1:]:
        sgn = 1 if f > 0 else -1
        inputs[("color", f, 0)] = (_shift(tgt, 3 * sgn, 1 * sgn) + 0.02 * torch.randn(tgt.shape, generator=gen)).clamp(0, 1)
    for f in opt.frame_ids:
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    for s in opt.scales:
        inputs[("K", s)], inputs[("inv_K", s)] = camera(kind, H // 2**s, W // 2**s, B)
    for f in opt.frame_ids[1:]:
        if kind == "nuscenes":
            inputs[("ts", f)] = torch.tensor([0.5, 1.0, 1.5])[torch.randint(0, 3, (B,), generator=gen)]
        else:
            inputs[("ts", f)] = torch.ones(B, dtype=torch.int64)
    return inputs


class SyntheticTriplets:
    """Iterable of `steps` batches; `pinned=True` returns pinned host tensors (the H2D copy then happens in
    Trainer.process_inputs, as in the reference), otherwise the batch is made resident on `device` once."""

    def __init__(self, opt, steps, device=None, seed=1234, pinned=False, distinct=2):
        self.opt, self.steps, self.device, self.pinned = opt, steps, device, pinned
        self.batches = []
        for i in range(distinct):
            b = make_batch(opt, seed + i)
            if pinned:
                b = {k: v.pin_memory() for k, v in b.items()}
            elif device is not None:
                b = {k: v.to(device) for k, v in b.items()}
            self.batches.append(b)

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield dict(self.batches[i % len(self.batches)])
