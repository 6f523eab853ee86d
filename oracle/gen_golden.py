"""ORACLE pinning (test infrastructure): execute the UNMODIFIED reference on CPU and record golden
input/output vectors under tests/golden/.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden [loss] [decoders] [step]

The reference ships no tests / golden vectors of its own, so these files are
the pin: tests/test_oracle_golden.py checks oracle/ against them on CPU, the -m gpu tests check
the CUDA path against them on the B200.  Nothing here is imported by the product.
"""
import os
import sys
import zlib

import numpy as np
import torch

from . import _refshim, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def key_str(k):
    if isinstance(k, tuple):
        return "|".join(str(x) for x in k)
    return str(k)


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


# ---------------------------------------------------------------------------------------------
# loss path: Trainer.generate_images_pred + Trainer.compute_losses (Trainer.py:215-423)
# ---------------------------------------------------------------------------------------------

LOSS_CASES = [
    # name, model, phase, B, H, W, ts_mode, ramp step (of 100), full tensors?
    dict(name="loss_dispinit_md2_32x64", model="monodepthv2", phase="disp_init", B=2, H=32, W=64, ts="ones", step=0, full=True, seed=11),
    dict(name="loss_motioninit_lite_32x64", model="litemono", phase="motion_init", B=2, H=32, W=64, ts="ones", step=20, full=True, seed=12),
    dict(name="loss_maskinit_lite_32x64", model="litemono", phase="mask_init", B=2, H=32, W=64, ts="ones", step=20, full=True, seed=13),
    dict(name="loss_finetune_md2_64x96", model="monodepthv2", phase="fine_tune", B=3, H=64, W=96, ts="float", step=50, full=False, seed=14),
    dict(name="loss_dispinit_lite_96x128", model="litemono", phase="disp_init", B=2, H=96, W=128, ts="ones", step=0, full=False, seed=15),
    # fine_tune with the RANSAC ground prior on (g_d_ground = 0.1); host RNG replaced by oracle.ground.SeededIndices
    dict(name="loss_finetune_ground_lite_64x96", model="litemono", phase="fine_tune", B=3, H=64, W=96, ts="ones", step=60, full=False,
         seed=16, ground=True),
]


def run_reference_loss(ns, case):
    argv = ["-d", "kitti", "--depth_model", case["model"], "--weights_init", "scratch", "-b", str(case["B"]),
            "--height", str(case["H"]), "--width", str(case["W"]), "--g_d_ground", "0.1" if case.get("ground") else "0.0"]
    tr = _refshim.make_reference_trainer(ns, argv, phase=case["phase"], step=case["step"], steps_per_epoch=100)
    scales = tr.opt.scales
    flow = case["phase"] != "disp_init"
    inputs, leaves = synth.make_loss_inputs(case["seed"], case["B"], case["H"], case["W"], scales, kind="kitti",
                                            flow=flow, ts_mode=case["ts"], all_scale_intrinsics=bool(case.get("ground")))
    tr.apply_img_resize(inputs)  # Trainer.py:729-734 builds ('color',0,s) for s>0
    noise = synth.automask_noise(case["seed"], case["B"], case["H"], case["W"], scales)

    outputs = {}
    grads_of = {}
    for k, v in leaves.items():
        v = v.clone().requires_grad_(True)
        grads_of[k] = v
        if k[0] == "motion_prob":
            s = k[1]
            for f in (-1, 1):  # networks/model.py:143-149: the same tensors for both frames
                outputs[("motion_prob", f, s)] = v
            m = torch.sigmoid(v)
            for f in (-1, 1):
                outputs[("motion_mask", f, s)] = m
        else:
            outputs[k] = v

    # inject the automask tie-break noise (Trainer.py:339 draws torch.randn internally)
    noise_queue = [noise[s] for s in scales]
    real_randn = torch.randn

    def fake_randn(*a, **kw):
        return noise_queue.pop(0).clone()

    real_choice = np.random.choice
    if case.get("ground"):
        from .ground import SeededIndices
        seeded = SeededIndices(case["seed"])
        np.random.choice = lambda a, size, replace=True: seeded(len(a), size)
    torch.randn = fake_randn
    try:
        tr.generate_images_pred(inputs, outputs)
        losses = tr.compute_losses(inputs, outputs)
    finally:
        torch.randn = real_randn
        np.random.choice = real_choice
    losses["loss"].backward()

    rec = {}
    # small cases carry their inputs; larger ones are re-synthesised in the tests from the seed
    # (oracle/synth.py is deterministic on CPU) and verified through these checksums.
    for k, v in list(inputs.items()) + [(("leaf",) + k, v) for k, v in leaves.items()]:
        a = _np(v).astype(np.float64)
        rec["chk:" + key_str(k)] = np.array([a.sum(), np.abs(a).sum()])
    if case["full"]:
        for k, v in inputs.items():
            rec["in:" + key_str(k)] = _np(v)
        for k, v in leaves.items():
            rec["leaf:" + key_str(k)] = _np(v)
        if case["phase"] == "disp_init":
            for s in scales:
                rec[f"noise:{s}"] = _np(noise[s])
    rec["meta:seed"] = np.asarray(case["seed"])
    rec["meta:ts_mode"] = np.asarray(case["ts"])
    rec["meta:shape"] = np.asarray([case["B"], case["H"], case["W"]])
    rec["meta:ground"] = np.asarray(1 if case.get("ground") else 0)
    for k, v in losses.items():
        rec["loss:" + key_str(k)] = np.float64(float(v))
    for k, v in grads_of.items():
        if v.grad is not None:
            rec["grad:" + key_str(k)] = _np(v.grad)
    keep = ["color", "sample", "residual_flow", "independ_flow", "depth", "sample_ego", "sample_complete"]
    for k, v in outputs.items():
        if isinstance(k, tuple) and k[0] in keep and case["full"]:
            rec["out:" + key_str(k)] = _np(v)
        if isinstance(k, str) and k.startswith("identity_selection"):
            rec["out:" + k] = _np(v)
    rec["meta:scales"] = np.asarray(scales)
    rec["meta:phase"] = np.asarray(case["phase"])
    rec["meta:step"] = np.asarray(case["step"])
    rec["meta:steps_per_epoch"] = np.asarray(100)
    return rec


def gen_loss(ns):
    for case in LOSS_CASES:
        rec = run_reference_loss(ns, case)
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, **rec)
        print(f"wrote {path}: {len(rec)} arrays, {os.path.getsize(path)/1024:.0f} KiB, loss={rec['loss:loss']:.6f}")


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    what = argv or ["loss", "decoders", "step"]
    ns = _refshim.load_reference()
    torch.set_num_threads(os.cpu_count())
    if "loss" in what:
        gen_loss(ns)
    if "decoders" in what:
        from . import gen_golden_nets
        gen_golden_nets.gen_decoders(ns)
    if "step" in what:
        from . import gen_golden_nets
        gen_golden_nets.gen_step(ns)


if __name__ == "__main__":
    main(sys.argv[1:])
