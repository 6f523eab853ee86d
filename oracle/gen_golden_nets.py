"""ORACLE pinning for the networks and the full step (run through `python -m oracle.gen_golden decoders step`).

Executes the UNMODIFIED reference decoders / Model / Trainer on CPU with key-addressed deterministic
weights (oracle.synth.fill_state) and records outputs + gradients.  Weights and inputs are re-created
from their seeds by the tests, so only results are stored.
"""
import json
import os

import numpy as np
import torch

from . import _refshim, synth
from .gen_golden import GOLDEN_DIR, _np, key_str


def chk(t):
    a = _np(t).astype(np.float64)
    return np.array([a.sum(), np.abs(a).sum(), float((a * a).sum())])


def seeded(shape, seed, scale=1.0):
    return scale * torch.randn(shape, generator=torch.Generator().manual_seed(seed))


DEC_B, DEC_H, DEC_W = 2, 64, 96
RES_CH = np.array([8, 8, 12, 16, 24])
LITE_CH = np.array([8, 12, 16])


def decoder_inputs(kind):
    """feature pyramids (seeded) for the small decoder configurations"""
    if kind == "md2":
        return [seeded((DEC_B, int(c), DEC_H >> (i + 1), DEC_W >> (i + 1)), 100 + i) for i, c in enumerate(RES_CH)]
    if kind == "lite":
        return [seeded((DEC_B, int(c), DEC_H >> (i + 2), DEC_W >> (i + 2)), 200 + i) for i, c in enumerate(LITE_CH)]
    if kind == "motion":
        feats = [seeded((DEC_B, 9, DEC_H, DEC_W), 300)]
        feats += [seeded((DEC_B, int(c), DEC_H >> (i + 1), DEC_W >> (i + 1)), 301 + i) for i, c in enumerate(RES_CH)]
        return feats
    if kind == "pose":
        return [seeded((DEC_B, int(RES_CH[-1]), DEC_H >> 5, DEC_W >> 5), 400)]
    raise ValueError(kind)


def record_module(rec, tag, module, inputs_requiring_grad, outputs):
    """objective = sum_k <out_k, G_k> with seeded G_k; records outputs and all gradients"""
    obj = 0
    for n, (k, v) in enumerate(sorted(outputs.items(), key=lambda kv: str(kv[0]))):
        G = seeded(v.shape, 900 + n)
        obj = obj + (v * G).sum()
        rec[f"{tag}:out:{key_str(k)}"] = _np(v)
    obj.backward()
    rec[f"{tag}:objective"] = np.float64(float(obj))
    for n, t in enumerate(inputs_requiring_grad):
        rec[f"{tag}:gin:{n}"] = _np(t.grad)
    for k, p in module.named_parameters():
        if k.startswith("net."):   # PoseDecoder aliases (pose_decoder.py:21)
            continue
        g = _np(p.grad)
        rec[f"{tag}:gchk:{k}"] = chk(g)
        if g.size <= 4096:
            rec[f"{tag}:gparam:{k}"] = g


def gen_decoders(ns):
    nets = ns.networks
    dd = __import__("networks.depth_decoder", fromlist=["x"])
    md = __import__("networks.motion_decoder", fromlist=["x"])
    pd = __import__("networks.pose_decoder", fromlist=["x"])
    rec = {}
    # Monodepth2 depth decoder
    m = dd.DepthDecoder(RES_CH, scales=range(4))
    synth.fill_state(m, 1)
    feats = [f.requires_grad_(True) for f in decoder_inputs("md2")]
    record_module(rec, "md2", m, feats, m(feats))
    # Lite-Mono depth decoder
    m = dd.LiteDepthDecoder(LITE_CH, scales=range(3))
    synth.fill_state(m, 2)
    feats = [f.requires_grad_(True) for f in decoder_inputs("lite")]
    record_module(rec, "lite", m, feats, m(feats))
    # motion decoders (flow / mask)
    for tag, od, seed in (("flow", 3, 3), ("mask", 1, 4)):
        m = md.MotionDecoder(RES_CH, [0, 1, 2, 3], num_input_images=3, inp_disp=False, out_dim=od)
        synth.fill_state(m, seed)
        feats = [f.requires_grad_(True) for f in decoder_inputs("motion")]
        ego = seeded((DEC_B, 6, 1, 1), 500, 0.01).requires_grad_(True)
        record_module(rec, tag, m, feats + [ego], m(feats, ego))
    # pose decoder + transformation_from_parameters
    m = pd.PoseDecoder(RES_CH, num_input_features=1, num_frames_to_predict_for=2)
    synth.fill_state(m, 5)
    feats = [f.requires_grad_(True) for f in decoder_inputs("pose")]
    aa, tr = m([feats])
    T = ns.networks.model.transformation_from_parameters(aa[:, 0] * 30, tr[:, 0] * 30, invert=True)
    record_module(rec, "pose", m, feats, {"axisangle": aa, "translation": tr, "T": T})
    path = os.path.join(GOLDEN_DIR, "nets_decoders.npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: {len(rec)} arrays, {os.path.getsize(path)/1024:.0f} KiB")


# ------------------------------------------------------------------------------------------------


def state_keys(ns):
    out = {}
    for dm in ("monodepthv2", "litemono"):
        opt = ns.options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", dm, "--weights_init", "scratch"])
        model = ns.networks.Model(opt)
        out[dm] = {name: [[k, list(v.shape)] for k, v in getattr(model, name).state_dict().items()] for name in sorted(model.module_names)}
    path = os.path.join(GOLDEN_DIR, "state_keys.json")
    json.dump(out, open(path, "w"))
    print(f"wrote {path}: {os.path.getsize(path)/1024:.0f} KiB")


def model_forward_golden(ns, depth_model, H, W, B, seed, name):
    """Model.forward in eval mode (deterministic BN / DropPath) with filled weights, all heads active."""
    opt = ns.options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", depth_model, "--weights_init", "scratch",
                                                 "-b", str(B), "--height", str(H), "--width", str(W)])
    torch.manual_seed(0)
    model = ns.networks.Model(opt)
    synth.fill_state(model, seed)
    model.set_eval()
    inputs, _ = synth.make_loss_inputs(seed, B, H, W, opt.scales, flow=False)
    with torch.no_grad():
        out = model(inputs)
    rec = {"meta:shape": np.asarray([B, H, W]), "meta:seed": np.asarray(seed), "meta:scales": np.asarray(opt.scales)}
    for k, v in out.items():
        if not torch.is_tensor(v):
            continue
        rec["chk:" + key_str(k)] = chk(v)
        if v.numel() <= 8192:
            rec["out:" + key_str(k)] = _np(v)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: {len(rec)} arrays, {os.path.getsize(path)/1024:.0f} KiB")


POSE_BIAS = [0.02, -0.03, 0.01, 0.18, 0.09, 0.05]   # x 0.01 in the decoder: ~2e-4 rad, (1.8, 0.9, 0.5) mm -> a (3.3, 1.7) px shift at 0.2 m


def step_golden(ns, automask=True):
    """BASELINE config 1: tiny_kitti 192x640 bs2 monodepthv2, reference Trainer, 1 step forward+loss
    (+ backward and one Adam step), phase disp_init, train mode, injected automask noise.
    automask=False records the WELL-CONDITIONED variant of the same step (step_config1_tiny_kitti_posed.npz): at random
    initial weights the predicted pose is the identity to ~1e-6, so every sampling coordinate sits ON the integer pixel
    lattice, where floor() -- and with it the bilinear coordinate gradient -- flips under last-ulp differences, and with
    auto-masking on almost every pixel is also a near-tie between the identity and the warped loss.  The variant switches
    auto-masking off and adds POSE_BIAS to the pose head's output bias (a few-pixel, non-integer shift), after which the
    gradients of all parameters can be compared tightly."""
    data_path = os.path.join(_refshim.REFERENCE_ROOT, "assets", "tiny_kitti") + "/"
    argv = ["-d", "kitti", "--depth_model", "monodepthv2", "--weights_init", "scratch", "-b", "2", "--data_path", data_path]
    tr = _refshim.make_reference_trainer(ns, argv, phase="disp_init", step=0, steps_per_epoch=100)
    synth.fill_state(tr.base_model, 21)
    if not automask:
        with torch.no_grad():
            tr.base_model.pose_dec.pose2.bias[:6] += torch.tensor(POSE_BIAS)
    tr.setup_phase("disp_init")   # optimiser over the freshly filled parameters
    tr.bool_automask = automask
    tr.set_train()
    files = ["2011_09_26/2011_09_26_drive_0001_sync 1 l"] * 2
    ds = tr.get_dataset(files, is_train=False, load_depth=False, load_mask=False)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, num_workers=0)
    inputs = next(iter(loader))
    rec = {}
    for f in (0, -1, 1):   # the three input frames as uint8 (ToTensor() divides by 255 exactly)
        img = inputs[("color", f, 0)][0]
        u8 = (img * 255).round().to(torch.uint8)
        assert torch.equal(u8.float() / 255, img)
        rec[f"img:{f}"] = _np(u8)
    rec["K"] = _np(inputs[("K", 0)][0])
    noise = synth.automask_noise(21, 2, 192, 640, tr.opt.scales)
    queue = [noise[s] for s in tr.opt.scales]
    real = torch.randn
    torch.randn = lambda *a, **k: queue.pop(0).clone()
    try:
        outputs, losses = tr.process_batch(inputs)
    finally:
        torch.randn = real
    losses["loss"].backward()
    for k, v in losses.items():
        rec["loss:" + k] = np.float64(float(v))
    for s in tr.opt.scales:
        rec[f"chk:disp|0|{s}"] = chk(outputs[("disp", 0, s)])
    for f in (-1, 1):
        rec[f"out:cam_T_cam|0|{f}"] = _np(outputs[("cam_T_cam", 0, f)])
    for mod in ("depth_enc", "depth_dec", "pose_enc", "pose_dec"):
        for k, p in getattr(tr.base_model, mod).named_parameters():
            if p.grad is not None and not k.startswith("net."):
                rec[f"gchk:{mod}.{k}"] = chk(p.grad)
    tr.optim["optimizer"].step()
    for mod in ("depth_dec", "pose_dec"):
        for k, p in getattr(tr.base_model, mod).named_parameters():
            if not k.startswith("net."):
                rec[f"pchk:{mod}.{k}"] = chk(p)
    if not automask:   # images / intrinsics live in the automask file
        rec = {k: v for k, v in rec.items() if k.startswith(("loss:", "gchk:", "chk:"))}
    path = os.path.join(GOLDEN_DIR, "step_config1_tiny_kitti.npz" if automask else "step_config1_tiny_kitti_posed.npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: {len(rec)} arrays, {os.path.getsize(path)/1024:.0f} KiB, loss={rec['loss:loss']:.6f}")


def gen_step(ns):
    state_keys(ns)
    model_forward_golden(ns, "monodepthv2", 64, 96, 2, 31, "model_fwd_md2_64x96")
    model_forward_golden(ns, "litemono", 64, 96, 2, 32, "model_fwd_lite_64x96")
    step_golden(ns)
    step_golden(ns, automask=False)
