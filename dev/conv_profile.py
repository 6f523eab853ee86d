"""Development: per-layer CUDA-event timing of every dd_conv_* call inside one training step."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dynamo-depth_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from dd_b200 import functional as Fn  # noqa: E402
from dd_b200 import synthetic  # noqa: E402
from Trainer import Trainer  # noqa: E402

records = collections.defaultdict(list)
orig_f, orig_b = Fn._ConvFn.forward, Fn._ConvFn.backward


def key_of(x0, x1, weight, ksize, pad_mode, act, up0):
    return (tuple(x0.shape), tuple(x1.shape) if x1 is not None else None, weight.shape[0], ksize, pad_mode, act, up0)


def fwd(ctx, x0, x1, weight, bias, residual, ksize, pad_mode, act, up0):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = orig_f(ctx, x0, x1, weight, bias, residual, ksize, pad_mode, act, up0)
    e.record()
    ctx.prof_key = key_of(x0, x1, weight, ksize, pad_mode, act, up0)
    records[("fwd",) + ctx.prof_key].append((s, e))
    return out


def bwd(ctx, go):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = orig_b(ctx, go)
    e.record()
    records[("bwd",) + ctx.prof_key].append((s, e))
    return out


Fn._ConvFn.forward = staticmethod(fwd)
Fn._ConvFn.backward = staticmethod(bwd)

torch.backends.cudnn.benchmark = True
opt = bench.make_opt(32)
tr = Trainer(opt)
tr.setup_phase("fine_tune")
tr.num_steps_per_epoch, tr.step = 100, 100
tr.set_train()
batches = synthetic.SyntheticTriplets(opt, steps=1, device=tr.device, distinct=1).batches
for _ in range(3):
    tr.train_step(dict(batches[0]))
records.clear()
tr.train_step(dict(batches[0]))
torch.cuda.synchronize()
rows = []
for k, evs in records.items():
    ms = sum(a.elapsed_time(b) for a, b in evs)
    x0, x1, cout, ks = k[1], k[2], k[3], k[4]
    cin = x0[1] + (x1[1] if x1 else 0)
    hw = (x1[2] * x1[3]) if x1 else (x0[2] * x0[3] * (4 if k[7] else 1))
    gf = 2.0 * x0[0] * cout * cin * ks * ks * hw / 1e9 * len(evs) * (2 if k[0] == "bwd" else 1)
    rows.append((ms, k[0], len(evs), cin, cout, ks, hw, gf / ms if ms > 0 else 0, k))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total conv time {tot:.1f} ms in {sum(r[2] for r in rows)} calls")
for ms, kind, n, cin, cout, ks, hw, tfs, k in rows[:45]:
    print(f"{ms:7.2f} ms {100*ms/tot:5.1f}%  {kind} x{n}  Cin {cin:4d} Cout {cout:4d} k{ks} hw {hw:6d} up{k[7]} pad{k[5]} act{k[6]}  {tfs:6.1f} TF/s")
