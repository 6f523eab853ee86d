"""ORACLE helper (test infrastructure): load the golden fixtures written by oracle/gen_golden.py."""
import os

import numpy as np
import torch

from . import synth
from .view_synthesis import LossConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def parse_key(s):
    parts = s.split("|")
    if len(parts) == 1:
        return s
    out = []
    for p in parts:
        try:
            out.append(int(p))
        except ValueError:
            out.append(p)
    return tuple(out)


class LossCase:
    """One golden case of the loss path: inputs, leaf tensors, injected noise, expected losses/grads/outputs."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.scales = [int(s) for s in z["meta:scales"]]
        self.phase = str(z["meta:phase"])
        self.step = int(z["meta:step"])
        self.steps_per_epoch = int(z["meta:steps_per_epoch"])
        self.B, self.H, self.W = (int(v) for v in z["meta:shape"])
        seed = int(z["meta:seed"])
        self.seed = seed
        self.ground = bool(int(z["meta:ground"])) if "meta:ground" in z.files else False
        ts_mode = str(z["meta:ts_mode"])
        have_inputs = any(k.startswith("in:") for k in z.files)
        if have_inputs:
            self.inputs = {parse_key(k[3:]): torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
            self.leaves = {parse_key(k[5:]): torch.from_numpy(z[k]) for k in z.files if k.startswith("leaf:")}
            self.noise = {int(k[6:]): torch.from_numpy(z[k]) for k in z.files if k.startswith("noise:")}
        else:
            self.inputs, self.leaves = synth.make_loss_inputs(seed, self.B, self.H, self.W, self.scales, kind="kitti",
                                                              flow=self.phase != "disp_init", ts_mode=ts_mode,
                                                              all_scale_intrinsics=self.ground)
            synth.add_color_pyramid(self.inputs, self.scales, self.H, self.W)
            self.noise = synth.automask_noise(seed, self.B, self.H, self.W, self.scales)
        # verify (re-)synthesised inputs against the recorded checksums
        for k in z.files:
            if not k.startswith("chk:"):
                continue
            key = parse_key(k[4:])
            t = self.leaves[key[1:]] if isinstance(key, tuple) and key[0] == "leaf" else self.inputs[key]
            a = t.numpy().astype(np.float64)
            got = np.array([a.sum(), np.abs(a).sum()])
            if not np.allclose(got, z[k], rtol=1e-9, atol=1e-9):
                raise AssertionError(f"{name}: input {key} does not reproduce (checksum {got} vs {z[k]})")
        if self.phase != "disp_init":
            self.noise = None
        self.losses = {k[5:]: float(z[k]) for k in z.files if k.startswith("loss:")}
        self.grads = {parse_key(k[5:]): torch.from_numpy(z[k]) for k in z.files if k.startswith("grad:")}
        self.outputs = {parse_key(k[4:]): torch.from_numpy(z[k]) for k in z.files if k.startswith("out:")}
        self.cfg = LossConfig(self.H, self.W, self.scales, phase=self.phase, g_d_ground=0.1 if self.ground else 0.0)

    def fresh_outputs(self, dtype=torch.float32, device="cpu"):
        """outputs dict as Model.forward would fill it; returns (outputs, leaf dict with requires_grad)."""
        outputs, leaves = {}, {}
        for k, v in self.leaves.items():
            v = v.to(device=device, dtype=dtype).clone().requires_grad_(True)
            leaves[k] = v
            if k[0] == "motion_prob":
                m = torch.sigmoid(v)
                for f in (-1, 1):
                    outputs[("motion_prob", f, k[1])] = v
                    outputs[("motion_mask", f, k[1])] = m
            else:
                outputs[k] = v
        return outputs, leaves

    def cast_inputs(self, dtype=torch.float32, device="cpu"):
        out = {}
        for k, v in self.inputs.items():
            if v.is_floating_point():
                out[k] = v.to(device=device, dtype=dtype)
            else:
                out[k] = v.to(device=device)
        return out


LOSS_CASE_NAMES = [
    "loss_dispinit_md2_32x64",
    "loss_motioninit_lite_32x64",
    "loss_maskinit_lite_32x64",
    "loss_finetune_md2_64x96",
    "loss_dispinit_lite_96x128",
    "loss_finetune_ground_lite_64x96",
]
