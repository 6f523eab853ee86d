"""Development audit: config-1 step (well-conditioned variant) -- relative error of ||grad||^2 per module against the
reference golden, for whatever kernel selection the DD_* environment variables pick.  python dev/config1_grad_audit.py"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dynamo-depth_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_networks_gpu as T  # noqa: E402
from oracle import nets_io  # noqa: E402

z0, opt, tr, inputs = T._config1_trainer_and_inputs(automask=False)
z = nets_io.load_npz("step_config1_tiny_kitti_posed")
outputs, losses = tr.process_batch(inputs)
losses["loss"].backward()
errs = T._grad_norm_errors(tr, z, "audit")
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DD_")) or "default"
for mod in ("depth_enc", "depth_dec", "pose_enc", "pose_dec"):
    v = [e for k, (e, n) in errs.items() if k.startswith(mod)]
    big = [e for k, (e, n) in errs.items() if k.startswith(mod) and n >= 1000]
    print(f"[{tag}] {mod}: median {statistics.median(v):.2e} max {max(v):.2e} max(numel>=1000) {max(big):.2e}")
worst = sorted(((e, n, k) for k, (e, n) in errs.items() if k.startswith("depth_dec")), reverse=True)[:6]
for e, n, k in worst:
    print(f"    {e:.2e} numel={n} {k}")
