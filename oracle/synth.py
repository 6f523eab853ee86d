"""ORACLE helper (test infrastructure): seeded synthetic inputs for the view-synthesis loss path.

Shapes/keys follow the reference's `inputs` / `outputs` dict contracts (SURVEY.md section 8b;
datasets/base_dataset.py:99-204, networks/model.py:58-149).  Pure torch-CPU; used by
oracle/gen_golden.py (to feed the real reference), by tests and by bench.py.
"""
import math

import numpy as np
import torch

# normalised intrinsics (fx/W, fy/H, cx/W, cy/H)
INTRINSICS = {
    "kitti": (0.58, 1.92, 0.5, 0.5),  # datasets/kitti_dataset.py:14-17
    "waymo": (1.06485, 1.59727, 0.49457, 0.49472),  # assets/tiny_waymo cam.json
    "nuscenes": (0.79151, 1.40713, 0.51017, 0.54612),
}


def intrinsics(kind, height, width, batch):
    fx, fy, cx, cy = INTRINSICS[kind]
    K = np.array([[fx * width, 0, cx * width, 0], [0, fy * height, cy * height, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                 dtype=np.float32)
    inv_K = np.linalg.pinv(K)  # datasets/base_dataset.py:160
    K = torch.from_numpy(K).unsqueeze(0).repeat(batch, 1, 1)
    inv_K = torch.from_numpy(inv_K.astype(np.float32)).unsqueeze(0).repeat(batch, 1, 1)
    return K.contiguous(), inv_K.contiguous()


def _box_blur(x, k):
    pad = k // 2
    x = torch.nn.functional.pad(x, (pad, pad, pad, pad), mode="reflect")
    return torch.nn.functional.avg_pool2d(x, k, 1)


def smooth_field(gen, shape, k=7):
    """Low-pass random field rescaled to [0,1] per tensor."""
    h, w = shape[-2:]
    k = min(k, 2 * (min(h, w) // 2) - 1)
    x = torch.rand(shape, generator=gen)
    if k >= 3:
        x = _box_blur(x, k)
    x = x - x.amin()
    return x / x.amax().clamp_min(1e-6)


def shift_image(img, dx, dy):
    """Integer translation with border replication (source frames = shifted target)."""
    B, C, H, W = img.shape
    ys = (torch.arange(H) + dy).clamp(0, H - 1)
    xs = (torch.arange(W) + dx).clamp(0, W - 1)
    return img[:, :, ys][:, :, :, xs]


def pose_matrix(axisangle, translation):
    """Rodrigues + translation, T = R^T * Trans(-t) (networks/layers.py:7-82, invert=True)."""
    B = axisangle.shape[0]
    angle = axisangle.norm(dim=1, keepdim=True)
    axis = axisangle / (angle + 1e-7)
    ca, sa = torch.cos(angle)[:, 0], torch.sin(angle)[:, 0]
    C = 1 - ca
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    R = torch.zeros(B, 4, 4)
    R[:, 0, 0] = x * x * C + ca
    R[:, 0, 1] = x * y * C - z * sa
    R[:, 0, 2] = z * x * C + y * sa
    R[:, 1, 0] = x * y * C + z * sa
    R[:, 1, 1] = y * y * C + ca
    R[:, 1, 2] = y * z * C - x * sa
    R[:, 2, 0] = z * x * C - y * sa
    R[:, 2, 1] = y * z * C + x * sa
    R[:, 2, 2] = z * z * C + ca
    R[:, 3, 3] = 1
    Tm = torch.eye(4).repeat(B, 1, 1)
    Tm[:, :3, 3] = -translation
    return torch.matmul(R.transpose(1, 2), Tm)


def make_loss_inputs(seed, batch, height, width, scales, kind="kitti", flow=True, ts_mode="ones",
                     frame_ids=(0, -1, 1), all_scale_intrinsics=False):
    """Returns (inputs, leaves): `inputs` as the data loader would deliver them (without the colour
    pyramid of scales>0, see `add_color_pyramid`), `leaves` = network-output tensors of the keys
    Model.forward produces (disp, cam_T_cam, complete_flow, motion_prob)."""
    gen = torch.Generator().manual_seed(seed)
    inputs, leaves = {}, {}
    tgt = smooth_field(gen, (batch, 3, height, width))
    inputs[("color", 0, 0)] = tgt
    for f in frame_ids[1:]:
        sgn = 1 if f > 0 else -1
        src = shift_image(tgt, 3 * sgn, 1 * sgn) + 0.02 * torch.randn(tgt.shape, generator=gen)
        inputs[("color", f, 0)] = src.clamp(0, 1)
    for f in frame_ids:
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    K, inv_K = intrinsics(kind, height, width, batch)
    inputs[("K", 0)], inputs[("inv_K", 0)] = K, inv_K
    if all_scale_intrinsics:   # datasets/base_dataset.py:154-163 provides K / inv_K for every pyramid level
        for s in scales:
            if s != 0:
                inputs[("K", s)], inputs[("inv_K", s)] = intrinsics(kind, height // 2**s, width // 2**s, batch)
    for f in frame_ids[1:]:
        if ts_mode == "ones":
            inputs[("ts", f)] = torch.ones(batch, dtype=torch.int64)
        else:  # fp32 time gaps (the reference's nuScenes loader yields float64; we feed fp32)
            inputs[("ts", f)] = torch.tensor([0.5, 1.0, 1.5])[torch.randint(0, 3, (batch,), generator=gen)]
    for s in scales:
        h, w = height // 2**s, width // 2**s
        leaves[("disp", 0, s)] = 0.004 + 0.25 * smooth_field(gen, (batch, 1, h, w), k=5) ** 2
    for f in frame_ids[1:]:
        sgn = 1.0 if f > 0 else -1.0
        aa = 0.01 * torch.randn(batch, 3, generator=gen)
        tr = 0.05 * torch.randn(batch, 3, generator=gen)
        tr[:, 2] += 0.25 * sgn
        leaves[("cam_T_cam", 0, f)] = pose_matrix(aa, tr)
    if flow:
        for s in scales:
            h, w = height // 2**s, width // 2**s
            base = 0.2 * (smooth_field(gen, (batch, 3, h, w), k=5) - 0.5)
            blob = (smooth_field(gen, (batch, 1, h, w), k=5) > 0.7).float()
            leaves[("complete_flow", 1, s)] = base * (0.3 + blob)
            leaves[("complete_flow", -1, s)] = -leaves[("complete_flow", 1, s)] + 0.01 * torch.randn(batch, 3, h, w, generator=gen)
            leaves[("motion_prob", s)] = 6.0 * (smooth_field(gen, (batch, 1, h, w), k=5) - 0.55)
    return inputs, leaves


def automask_noise(seed, batch, height, width, scales, nframes=2):
    gen = torch.Generator().manual_seed(seed + 7919)
    return {s: torch.randn(batch, nframes, height, width, generator=gen) for s in scales}


def add_color_pyramid(inputs, scales, height, width):
    """('color',0,s) for s>0: chained bicubic antialias x1/2 + clamp (Trainer.py:80,729-734).
    Same third-party call the reference makes (torchvision Resize); moving it on-device is SURVEY 8f-2."""
    import torchvision.transforms as T

    for s in scales:
        if s != 0:
            rs = T.Resize((height // 2**s, width // 2**s), interpolation=T.InterpolationMode.BICUBIC, antialias=True)
            inputs[("color", 0, s)] = torch.clamp(rs(inputs[("color", 0, s - 1)]), 0, 1)
    return inputs


def fill_state(module_or_state, seed):
    """Deterministic, key-addressed parameter fill shared by the reference, the oracle and the product
    (no weight files ship): every tensor is drawn from its own generator seeded by crc32(key)."""
    import zlib

    sd = module_or_state.state_dict() if hasattr(module_or_state, "state_dict") else module_or_state
    out = {}
    for key in sorted(sd.keys()):
        t = sd[key]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**31 - 1))
        leaf = key.split(".")[-1]
        if not torch.is_tensor(t):
            out[key] = t
        elif leaf == "num_batches_tracked":
            out[key] = torch.zeros_like(t)
        elif leaf == "running_mean":
            out[key] = 0.1 * torch.randn(t.shape, generator=g)
        elif leaf == "running_var":
            out[key] = 0.8 + 0.4 * torch.rand(t.shape, generator=g)
        elif leaf in ("gamma", "gamma_xca"):
            out[key] = 0.2 * torch.randn(t.shape, generator=g)
        elif leaf == "temperature":
            out[key] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif leaf == "bias":
            out[key] = 0.05 * torch.randn(t.shape, generator=g)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            out[key] = torch.randn(t.shape, generator=g) * (1.6 / fan_in) ** 0.5
        else:  # 1-D scale of a normalisation layer
            out[key] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        if torch.is_tensor(out[key]):
            out[key] = out[key].to(t.dtype)
    if hasattr(module_or_state, "load_state_dict"):
        module_or_state.load_state_dict(out)
    return out
