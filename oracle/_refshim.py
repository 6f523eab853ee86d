"""Import shim for the UNMODIFIED reference tree (test infrastructure only).

The reference (/root/reference, YihongSun/Dynamo-Depth @227a5d9) is pure Python/PyTorch
but imports packages that are absent in this image (timm, imageio, matplotlib, skimage)
and asserts a visible CUDA device (Trainer.py:32).  This module registers the minimal
stand-ins and returns the reference's own modules so `oracle/gen_golden.py` can execute
the reference's code paths on CPU and record golden vectors.

Nothing in the product (dynamo-depth_b200/) or in the GPU tests imports this file; the
reference tree does not exist on the GPU box.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """$DD_REFERENCE_ROOT, else the read-only tree of the build container, else the unmodified install that
    __graft_entry__.build() puts under baseline/_ref (git-ignored; the only copy that exists on the GPU box)."""
    for cand in (os.environ.get("DD_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "Trainer.py")):
            return cand
    return os.environ.get("DD_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


class _DropPath(nn.Module):
    """timm==0.6.13 DropPath semantics: per-sample Bernoulli(keep)/keep, identity in eval."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Trainer.py"))


def load_reference():
    """Returns a namespace with the reference's `tools`, `utils`, `options`, `networks`, `Trainer` modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    if "timm" not in sys.modules:
        timm = _stub("timm")
        models = _stub("timm.models")
        layers = _stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=nn.init.trunc_normal_)
        timm.models = models
        models.layers = layers
    for name in ("imageio", "matplotlib", "matplotlib.cm", "skimage", "skimage.transform", "wandb"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "cm"):
        sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    if "skimage" in sys.modules and not hasattr(sys.modules["skimage"], "transform"):
        sys.modules["skimage"].transform = sys.modules["skimage.transform"]

    import PIL.Image as pil_image

    if not hasattr(pil_image, "ANTIALIAS"):
        pil_image.ANTIALIAS = pil_image.LANCZOS

    # Trainer.py:32 asserts cuda_id < device_count(); pretend one device so the CPU path constructs.
    if torch.cuda.device_count() == 0:
        torch.cuda.device_count = lambda: 1

    # The reference uses top-level module names (tools, utils, options, networks, datasets, Trainer)
    # that collide with the product's drop-in names; make sure the reference wins in THIS process.
    for name in ("tools", "utils", "options", "networks", "datasets", "Trainer"):
        for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            del sys.modules[key]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ns = types.SimpleNamespace()
        ns.tools = importlib.import_module("tools")
        ns.utils = importlib.import_module("utils")
        ns.options = importlib.import_module("options")
        ns.networks = importlib.import_module("networks")
        ns.Trainer = importlib.import_module("Trainer")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    for mod in (ns.tools, ns.utils, ns.options, ns.networks, ns.Trainer):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
    return ns


def make_reference_trainer(ns, argv, phase="disp_init", step=0, steps_per_epoch=100):
    """Build the reference Trainer on CPU without touching data loaders / wandb (never call train()/val())."""
    opt = ns.options.DynamoOptions().parse(args=argv)
    opt.local_world_size = 1
    opt.ddp = False
    # save_opt() writes opt.json under log_dir; keep it out of the repo
    opt.log_dir = os.environ.get("DD_REF_LOGDIR", "/tmp/dd_ref_logs")
    tr = ns.Trainer.Trainer(opt)
    tr.setup_phase(phase)
    tr.step = step
    tr.num_steps_per_epoch = steps_per_epoch
    tr.bool_automask = phase == "disp_init"
    return tr
