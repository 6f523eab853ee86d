"""Step-time experiments on the PyTorch-side encoders (not product defaults): channels_last ResNet encoders, TF32 matmul
for the Lite-Mono linear layers.  python dev/step_experiments.py [--steps 8]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dynamo-depth_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def run(tag, steps, channels_last=False, tf32=False, phase="fine_tune"):
    from Trainer import Trainer
    from dd_b200 import synthetic
    torch.backends.cudnn.benchmark = True
    opt = bench.make_opt(bench.BATCH, 0)
    torch.manual_seed(1234)
    tr = Trainer(opt)
    tr.setup_phase(phase)
    tr.bool_automask = phase == "disp_init"
    tr.num_steps_per_epoch, tr.step = 100, 100
    tr.set_train()
    if channels_last:   # Lite-Mono encoder in NHWC as well (the ResNet encoders already are by default)
        tr.model.depth_enc.to(memory_format=torch.channels_last)
        orig = tr.model.depth_enc.forward
        tr.model.depth_enc.forward = lambda x: orig(x.contiguous(memory_format=torch.channels_last))
    from networks.depth_encoder import EncoderLinear
    EncoderLinear.tf32 = tf32
    batches = synthetic.SyntheticTriplets(opt, steps=1, device=tr.device, seed=1234, distinct=2).batches
    for i in range(3):
        tr.train_step(dict(batches[i % 2]))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        out, losses = tr.train_step(dict(batches[i % 2]))
        del out
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    print(f"{tag:40s} {ms:8.2f} ms/step  {bench.BATCH / ms * 1e3:7.1f} triplets/s  loss {float(losses['loss']):.6f}", flush=True)
    del tr
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    run("baseline fine_tune (warm-up run)", a.steps)
    run("baseline fine_tune", a.steps)
    run("lite-mono encoder channels_last", a.steps, channels_last=True)
    run("scoped tf32 lite-mono linear layers", a.steps, tf32=True)
