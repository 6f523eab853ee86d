"""GPU parity of the STAND-ALONE module surface (the classes north_star names: BackprojectDepth, Project3D, SSIM,
compute_smooth_loss -- `tools.py:167-326` in the reference, also exported from `networks.layers`): each goes through its
own C-ABI kernel (dd_backproject_*, dd_project_*, dd_ssim_*, dd_smooth_*) and is compared, values and gradients,
with the golden recorded from the unmodified reference tools.py (oracle/gen_golden_tools.py) at 1e-4 relative fp32,
and with the CPU oracle on a second, larger, non-square case."""
import os

import numpy as np
import pytest
import torch

from oracle import view_synthesis as vs
from oracle.compare import check_rel

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tools_standalone.npz"))
    return z, {k[3:]: torch.from_numpy(z[k]).cuda() for k in z.files if k.startswith("in:")}


def test_backproject_and_project3d_match_reference_tools():
    import tools
    z, d = _golden()
    B, _, H, W = d["depth"].shape
    depth = d["depth"].clone().requires_grad_(True)
    cam = tools.BackprojectDepth(B, H, W)(depth, d["inv_K"])
    (cam * d["ct_cam"]).sum().backward()
    check_rel(cam, z["backproject:out"], 1e-5, what="backproject out")
    check_rel(depth.grad, z["backproject:g_depth"], 1e-5, what="backproject grad depth")
    for tag, T in (("project_T", d["T"]), ("project_noT", None)):
        pts = cam.detach().clone().requires_grad_(True)
        Tt = T.clone().requires_grad_(True) if T is not None else None
        pix, ego = tools.Project3D(B, H, W)(pts, d["K"], Tt)
        assert pix.shape == (B, H, W, 2) and ego.shape == (B, 3, H * W)
        ((pix * d["ct_pix"]).sum() + (ego * d["ct_ego"]).sum()).backward()
        check_rel(pix, z[f"{tag}:pix"], 1e-5, what=f"{tag} pix")
        check_rel(ego, z[f"{tag}:ego"], 1e-4, abs_tol=1e-6, what=f"{tag} ego")
        check_rel(pts.grad, z[f"{tag}:g_points"], 1e-4, what=f"{tag} grad points")
        if Tt is not None:
            check_rel(Tt.grad, z[f"{tag}:g_T"], 1e-4, what=f"{tag} grad T")


def test_layers_are_exported_from_networks_layers_too():
    import tools
    from networks import layers
    for name in ("BackprojectDepth", "Project3D", "SSIM", "compute_smooth_loss"):
        assert getattr(layers, name) is getattr(tools, name)


def test_ssim_matches_reference_tools():
    import tools
    z, d = _golden()
    x, y = d["x"].clone().requires_grad_(True), d["y"].clone().requires_grad_(True)
    s = tools.SSIM()(x, y)
    (s * d["ct_ssim"]).sum().backward()
    check_rel(s, z["ssim:out"], 1e-4, what="ssim out")
    check_rel(x.grad, z["ssim:g_x"], 1e-4, what="ssim grad x")
    check_rel(y.grad, z["ssim:g_y"], 1e-4, what="ssim grad y")


def test_compute_smooth_loss_matches_reference_tools():
    import tools
    z, d = _golden()
    for tag, inp, img in (("smooth1", d["smooth_inp1"], d["smooth_img"]), ("smooth3", d["smooth_inp3"], d["smooth_img"]),
                          ("smooth_noimg", d["smooth_inp3"], None)):
        t = inp.clone().requires_grad_(True)
        v = tools.compute_smooth_loss(t, img)
        v.backward()
        check_rel(v, z[f"{tag}:out"], 1e-5, what=f"{tag} value")
        check_rel(t.grad, z[f"{tag}:g_inp"], 1e-4, what=f"{tag} grad")


def test_standalone_layers_vs_oracle_larger_case():
    """Non-square 3-image case with clamped SSIM regions (identical / inverted patches) and points behind the camera."""
    import tools
    g = torch.Generator().manual_seed(5)
    B, H, W = 3, 40, 72
    x, y = torch.rand(B, 3, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)
    y[:, :, :8] = x[:, :, :8]                 # SSIM = 1 -> loss 0 (clamp edge)
    y[:, :, 8:16] = 1 - x[:, :, 8:16]         # anti-correlated -> upper clamp region
    ct = torch.randn(B, 3, H, W, generator=g)
    xo, yo = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    (vs.ssim(xo, yo) * ct).sum().backward()
    xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    s = tools.SSIM()(xc, yc)
    (s * ct.cuda()).sum().backward()
    check_rel(s, vs.ssim(x, y), 1e-4, abs_tol=1e-6, what="ssim out (clamped regions)")
    check_rel(xc.grad, xo.grad, 2e-4, abs_tol=1e-6, what="ssim grad x (clamped regions)")
    K = torch.tensor([[1.06 * W, 0, 0.49 * W, 0], [0, 1.6 * H, 0.49 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).repeat(B, 1, 1)
    inv_K = torch.linalg.pinv(K)
    depth = 0.5 + 30 * torch.rand(B, 1, H, W, generator=g)
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, :3, 3] = torch.tensor([0.3, -0.1, -1.5])          # z shift: points closer than 1.5 m end up behind the camera
    T[:, :3, :3] += 0.05 * torch.randn(B, 3, 3, generator=g)
    ctp, cte = torch.randn(B, H, W, 2, generator=g), torch.randn(B, 3, H * W, generator=g)
    do, To = depth.clone().requires_grad_(True), T.clone().requires_grad_(True)
    pix_o, ego_o = vs.project(vs.backproject(do, inv_K), K, To, H, W)
    ((pix_o * ctp).sum() + (ego_o * cte).sum()).backward()
    dc, Tc = depth.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
    pix, ego = tools.Project3D(B, H, W)(tools.BackprojectDepth(B, H, W)(dc, inv_K.cuda()), K.cuda(), Tc)
    ((pix * ctp.cuda()).sum() + (ego * cte.cuda()).sum()).backward()
    # 1/(z + eps) amplifies rounding near z = 0: compare where the reference's own result is well conditioned
    ok = (pix_o.abs() < 50).all(-1)
    check_rel(pix.detach().cpu()[ok], pix_o.detach()[ok], 1e-4, abs_tol=1e-4, what="project pix (well conditioned)")
    check_rel(ego, ego_o, 1e-4, abs_tol=1e-6, what="project ego")
    check_rel(Tc.grad, To.grad, 5e-4, what="project grad T (incl. near-singular points)")
    check_rel(dc.grad, do.grad, 5e-4, what="backproject+project grad depth")
