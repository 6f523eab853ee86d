"""ORACLE pinning (test infrastructure): golden vectors of the STAND-ALONE geometry / photometric layers, recorded by
executing the unmodified reference `tools.py` (BackprojectDepth :167-197, Project3D :200-224, SSIM :227-257,
compute_smooth_loss :311-326, disp_to_depth :291-298) on CPU in the build container.

    python -m oracle.gen_golden_tools          ->  tests/golden/tools_standalone.npz

Each layer is evaluated on small seeded inputs, a fixed random cotangent is back-propagated, and inputs, outputs and
input gradients are stored.  tests/test_oracle_golden.py pins oracle/view_synthesis.py to the file on CPU,
tests/test_tools_gpu.py checks dd_backproject_* / dd_project_* / dd_ssim_* / dd_smooth_* against it on the B200.
"""
import importlib.util
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tools_standalone.npz")
B, H, W = 2, 24, 40


def make_inputs(seed=77):
    """Seeded inputs shared by the generator and (through the stored arrays) by the tests."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    K = K.unsqueeze(0).repeat(B, 1, 1)
    K[1, 0, 0] *= 1.1
    inv_K = torch.from_numpy(np.stack([np.linalg.pinv(k.numpy()) for k in K])).float()
    T = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
    T[:, :3, :3] += 0.02 * n(B, 3, 3)
    T[:, :3, 3] = 0.2 * n(B, 3)
    d = {"depth": 1.0 + 20.0 * r(B, 1, H, W), "K": K, "inv_K": inv_K, "T": T,
         "x": r(B, 3, H, W), "y": r(B, 3, H, W),
         "smooth_inp1": r(B, 1, H, W), "smooth_inp3": n(B, 3, H, W) * 0.1, "smooth_img": r(B, 3, H, W),
         "disp": r(B, 1, H, W)}
    d["ct_cam"] = n(B, 4, H * W)
    d["ct_pix"], d["ct_ego"] = n(B, H, W, 2), n(B, 3, H * W)
    d["ct_ssim"] = n(B, 3, H, W)
    return d


def main():
    spec = importlib.util.spec_from_file_location("ref_tools", os.path.join(os.environ.get("DD_REFERENCE_ROOT", "/root/reference"), "tools.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    d = make_inputs()
    out = {f"in:{k}": v.numpy() for k, v in d.items()}

    depth = d["depth"].clone().requires_grad_(True)
    cam = ref.BackprojectDepth(B, H, W)(depth, d["inv_K"])
    (cam * d["ct_cam"]).sum().backward()
    out["backproject:out"], out["backproject:g_depth"] = cam.detach().numpy(), depth.grad.numpy()

    for tag, T in (("project_T", d["T"]), ("project_noT", None)):
        pts = cam.detach().clone().requires_grad_(True)
        Tt = T.clone().requires_grad_(True) if T is not None else None
        pix, ego = ref.Project3D(B, H, W)(pts, d["K"], Tt)
        ((pix * d["ct_pix"]).sum() + (ego * d["ct_ego"]).sum()).backward()
        out[f"{tag}:pix"], out[f"{tag}:ego"], out[f"{tag}:g_points"] = pix.detach().numpy(), ego.detach().numpy(), pts.grad.numpy()
        if Tt is not None:
            out[f"{tag}:g_T"] = Tt.grad.numpy()

    x, y = d["x"].clone().requires_grad_(True), d["y"].clone().requires_grad_(True)
    s = ref.SSIM()(x, y)
    (s * d["ct_ssim"]).sum().backward()
    out["ssim:out"], out["ssim:g_x"], out["ssim:g_y"] = s.detach().numpy(), x.grad.numpy(), y.grad.numpy()

    for tag, inp, img in (("smooth1", d["smooth_inp1"], d["smooth_img"]), ("smooth3", d["smooth_inp3"], d["smooth_img"]),
                          ("smooth_noimg", d["smooth_inp3"], None)):
        t = inp.clone().requires_grad_(True)
        v = ref.compute_smooth_loss(t, img)
        v.backward()
        out[f"{tag}:out"], out[f"{tag}:g_inp"] = v.detach().numpy(), t.grad.numpy()

    scaled, dep = ref.disp_to_depth(d["disp"], 0.1, 100.0)
    out["disp_to_depth:scaled"], out["disp_to_depth:depth"] = scaled.numpy(), dep.numpy()
    out["depth_to_disp:out"] = ref.depth_to_disp(dep, 0.1, 100.0).numpy()
    np.savez_compressed(GOLDEN, **out)
    print(f"wrote {GOLDEN}: {len(out)} arrays, {os.path.getsize(GOLDEN) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
