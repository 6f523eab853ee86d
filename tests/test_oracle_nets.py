"""Pins oracle/networks.py (plain-PyTorch restatement of decoders / model wiring) and the product's
host-side module definitions (state_dict keys, encoders) against goldens recorded from the reference.
CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import networks as on
from oracle import nets_io, synth
from oracle.golden_io import GOLDEN_DIR


def _state(tag):
    return {k: v.detach().clone() for k, v in nets_io.product_decoder(tag).state_dict().items()}


def _run_oracle(tag):
    sd = {k: v.requires_grad_(True) for k, v in _state(tag).items()}
    if tag in ("md2", "lite"):
        feats = [f.requires_grad_(True) for f in nets_io.decoder_inputs(tag)]
        fn = on.depth_decoder_md2 if tag == "md2" else on.depth_decoder_lite
        out = fn(feats, sd, range(4) if tag == "md2" else range(3))
        ins = feats
    elif tag in ("flow", "mask"):
        feats = [f.requires_grad_(True) for f in nets_io.decoder_inputs("motion")]
        ego = nets_io.seeded((nets_io.DEC_B, 6, 1, 1), 500, 0.01).requires_grad_(True)
        out = on.motion_decoder(feats, ego, sd, [0, 1, 2, 3], 3 if tag == "flow" else 1)
        ins = feats + [ego]
    else:
        feats = [f.requires_grad_(True) for f in nets_io.decoder_inputs("pose")]
        aa, tr = on.pose_decoder(feats[0], sd)
        T = on.transformation_from_parameters(aa[:, 0] * 30, tr[:, 0] * 30, invert=True)
        out = {"axisangle": aa, "translation": tr, "T": T}
        ins = feats
    nets_io.objective(out).backward()
    return out, ins, sd


@pytest.mark.parametrize("tag", ["md2", "lite", "flow", "mask", "pose"])
def test_oracle_decoders_match_reference(tag):
    z = nets_io.load_npz("nets_decoders")
    out, ins, sd = _run_oracle(tag)
    ref = nets_io.golden_outputs(z, tag)
    assert set(map(str, ref)) == set(map(str, out))
    for k, v in ref.items():
        assert (out[k].detach() - v).abs().max().item() <= 2e-5 * (1 + v.abs().max().item()), (tag, k)
    for n, t in enumerate(ins):
        g = torch.from_numpy(z[f"{tag}:gin:{n}"])
        assert (t.grad - g).abs().max().item() <= 1e-4 * (g.abs().max().item() + 1e-9), (tag, "gin", n)
    for k in z.files:
        if k.startswith(f"{tag}:gchk:"):
            name = k[len(f"{tag}:gchk:"):]
            got = nets_io.chk(sd[name].grad)
            assert np.allclose(got, z[k], rtol=2e-4, atol=1e-6), (tag, name, got, z[k])


def test_product_state_dict_keys_match_reference():
    import networks
    import options

    ref = json.load(open(os.path.join(GOLDEN_DIR, "state_keys.json")))
    for dm, mods in ref.items():
        opt = options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", dm, "--weights_init", "scratch"])
        model = networks.Model(opt)
        assert sorted(model.module_names) == sorted(mods.keys())
        for name, entries in mods.items():
            own = getattr(model, name).state_dict()
            assert [k for k, _ in entries] == list(own.keys()), (dm, name)
            for k, shape in entries:
                assert list(own[k].shape) == shape, (dm, name, k)


def _oracle_model_from_product(depth_model, H, W, B, seed):
    import networks
    import options

    opt = options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", depth_model, "--weights_init", "scratch", "-b", str(B),
                                              "--height", str(H), "--width", str(W)])
    model = networks.Model(opt)
    synth.fill_state(model, seed)
    model.set_eval()
    states = {n: {k: v for k, v in getattr(model, n).state_dict().items() if not k.startswith("net.")}
              for n in ("depth_dec", "pose_dec", "motion_dec", "motion_mask")}
    om = on.OracleModel(depth_model, opt.scales, opt.frame_ids, model.depth_enc, model.pose_enc, model.motion_enc, states)
    om.eval()
    return opt, om


@pytest.mark.parametrize("dm,name,seed", [("monodepthv2", "model_fwd_md2_64x96", 31), ("litemono", "model_fwd_lite_64x96", 32)])
def test_model_forward_matches_reference(dm, name, seed):
    """product encoders (PyTorch, CPU) + oracle decoders / wiring reproduce the reference Model.forward"""
    z = nets_io.load_npz(name)
    B, H, W = (int(v) for v in z["meta:shape"])
    opt, om = _oracle_model_from_product(dm, H, W, B, seed)
    inputs, _ = synth.make_loss_inputs(seed, B, H, W, opt.scales, flow=False)
    with torch.no_grad():
        out = om(inputs)
    n = 0
    for k in z.files:
        if not k.startswith("chk:"):
            continue
        from oracle.golden_io import parse_key
        key = parse_key(k[4:])
        if key not in out:
            assert key[0] in ("pose_feats", "motion_feats"), key
            continue
        got = nets_io.chk(out[key])
        assert np.allclose(got, z[k], rtol=3e-4, atol=1e-5), (key, got, z[k])
        n += 1
    assert n >= 10
