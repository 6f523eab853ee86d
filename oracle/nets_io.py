"""ORACLE helper (test infrastructure): rebuild the seeded decoder / model configurations of
oracle/gen_golden_nets.py and load their golden results."""
import os

import numpy as np
import torch

from . import synth
from .gen_golden_nets import DEC_B, DEC_H, DEC_W, LITE_CH, RES_CH, decoder_inputs, seeded  # noqa: F401
from .golden_io import GOLDEN_DIR, parse_key


def load_npz(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def golden_outputs(z, tag):
    pre = f"{tag}:out:"
    return {parse_key(k[len(pre):]): torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}


def objective(outputs):
    """the seeded linear objective the golden gradients belong to"""
    obj = 0
    for n, (k, v) in enumerate(sorted(outputs.items(), key=lambda kv: str(kv[0]))):
        obj = obj + (v * seeded(v.shape, 900 + n).to(v.device)).sum()
    return obj


def chk(t):
    a = t.detach().double().cpu().numpy()
    return np.array([a.sum(), np.abs(a).sum(), float((a * a).sum())])


DECODER_SEEDS = {"md2": 1, "lite": 2, "flow": 3, "mask": 4, "pose": 5}


def product_decoder(tag):
    """The product's decoder module (constructed on CPU, weights filled by key) for a golden tag."""
    import networks.depth_decoder as dd
    import networks.motion_decoder as md
    import networks.pose_decoder as pd

    if tag == "md2":
        m = dd.DepthDecoder(RES_CH, scales=range(4))
    elif tag == "lite":
        m = dd.LiteDepthDecoder(LITE_CH, scales=range(3))
    elif tag in ("flow", "mask"):
        m = md.MotionDecoder(RES_CH, [0, 1, 2, 3], num_input_images=3, inp_disp=False, out_dim=3 if tag == "flow" else 1)
    else:
        m = pd.PoseDecoder(RES_CH, num_input_features=1, num_frames_to_predict_for=2)
    synth.fill_state(m, DECODER_SEEDS[tag])
    return m
