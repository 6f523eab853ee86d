"""ORACLE helper (test infrastructure): tolerant comparison for the view-synthesis path.

The path contains two discrete decisions whose outcome can flip under last-ulp differences of fp32
arithmetic (SURVEY.md section 7 "bit-level habits"): the per-pixel argmin over candidate losses
(Trainer.py:347) and the floor() of the sampling coordinate (Trainer.py:281).  A flip changes a 3x3
(SSIM window) patch of per-pixel gradients discontinuously, so per-pixel comparisons allow a small
fraction of outliers while integrated quantities (losses, pose gradients, L2 norms) stay tight.
"""
import torch

from . import parity_log

CURRENT_CASE = "unnamed"      # set per test by tests/conftest.py (pytest node name) so that every comparison is logged


def robust_report(got, ref, rtol=1e-4):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = ref.abs().max().item() + 1e-30
    d = (got - ref).abs()
    return {
        "max_err": d.max().item(),
        "scale": scale,
        "outlier_frac": (d > rtol * scale).double().mean().item(),
        "rel_l2": ((got - ref).norm() / (ref.norm() + 1e-30)).item(),
    }


def assert_close_robust(got, ref, rtol=1e-4, max_outlier_frac=2e-3, max_rel_l2=2e-2, what=""):
    """A single argmin / floor flip perturbs a 3x3 full-resolution patch, i.e. up to ~16 samples of a
    coarse pyramid level after the bilinear transpose; small maps therefore get an absolute allowance of
    32 samples (two flips) on top of the relative one (0 stays 0: used for integrated quantities)."""
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    r = robust_report(got, ref, rtol)
    if max_outlier_frac > 0:
        max_outlier_frac = max(max_outlier_frac, 32.0 / ref.numel())
    parity_log.record(CURRENT_CASE, what, max_abs_rel=r["max_err"] / r["scale"], rel_l2=r["rel_l2"], outlier_frac=r["outlier_frac"],
                      rtol=rtol, limit_rel_l2=max_rel_l2, limit_outlier_frac=max_outlier_frac)
    assert r["outlier_frac"] <= max_outlier_frac, (what, r)
    assert r["rel_l2"] <= max_rel_l2, (what, r)
    return r


def check_rel(got, ref, rel=1e-4, abs_tol=0.0, what=""):
    """Scalar (or max-abs over a tensor) relative comparison that also logs the measured error."""
    if torch.is_tensor(got) or torch.is_tensor(ref):
        g = torch.as_tensor(got).detach().double().cpu()
        r = torch.as_tensor(ref).detach().double().cpu()
        err, scale = (g - r).abs().max().item(), r.abs().max().item()
    else:
        err, scale = abs(float(got) - float(ref)), abs(float(ref))
    parity_log.record(CURRENT_CASE, what, max_abs_rel=err / (scale + 1e-30), limit=rel)
    assert err <= rel * scale + abs_tol, (what, err, scale, rel)
    return err / (scale + 1e-30)
