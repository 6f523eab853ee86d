"""GPU parity of Trainer.generate_images_pred + Trainer.compute_losses (the product's loss assembly on top of
the fused kernels) against the goldens recorded from the reference's own Trainer, all phases, incl. the
RANSAC ground prior of fine_tune.  Tolerance 1e-4 relative on every loss entry."""
import pytest
import torch

from oracle.compare import assert_close_robust, check_rel, grad_bounds
from oracle.golden_io import LOSS_CASE_NAMES, LossCase

pytestmark = pytest.mark.gpu


def _trainer_for(case):
    import options
    from Trainer import Trainer

    dm = "monodepthv2" if len(case.scales) == 4 else "litemono"
    opt = options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", dm, "--weights_init", "scratch", "-b", str(case.B),
                                              "--height", str(case.H), "--width", str(case.W),
                                              "--g_d_ground", "0.1" if case.ground else "0.0"])
    opt.ddp = False
    tr = Trainer(opt)
    tr.setup_phase(case.phase)
    tr.bool_automask = case.phase == "disp_init"
    tr.step, tr.num_steps_per_epoch = case.step, case.steps_per_epoch
    if case.noise is not None:
        tr.automask_noise = case.noise
    if case.ground:
        from oracle.ground import SeededIndices
        tr.gplane.rand_index_fn = SeededIndices(case.seed)
    return tr


@pytest.mark.parametrize("name", LOSS_CASE_NAMES)
def test_trainer_losses_match_reference(name):
    case = LossCase(name)
    tr = _trainer_for(case)
    inputs = case.cast_inputs(device="cuda")
    outputs, leaves = case.fresh_outputs(device="cuda")
    tr.generate_images_pred(inputs, outputs)
    losses = tr.compute_losses(inputs, outputs)
    losses["loss"].backward()
    torch.cuda.synchronize()
    assert set(case.losses) <= set(losses), set(case.losses) - set(losses)
    for k, ref in case.losses.items():
        got = losses[k]
        got = float(got.detach()) if torch.is_tensor(got) else float(got)
        check_rel(got, ref, 1e-4, abs_tol=1e-7, what=k)
    for k, ref in case.grads.items():
        got = leaves[k].grad
        assert got is not None, k
        assert_close_robust(got.cpu(), ref, what=k, **grad_bounds(name, k))


def test_materialised_outputs_through_trainer():
    case = LossCase("loss_maskinit_lite_32x64")
    tr = _trainer_for(case)
    tr.materialise_outputs = True
    inputs = case.cast_inputs(device="cuda")
    outputs, _ = case.fresh_outputs(device="cuda")
    tr.generate_images_pred(inputs, outputs)
    for k, ref in case.outputs.items():
        if isinstance(k, tuple) and k[0] in ("color", "sample", "residual_flow", "independ_flow", "depth"):
            assert k in outputs, k
            assert_close_robust(outputs[k].detach().cpu(), ref, rtol=3e-4, max_rel_l2=1e-3, what=k)


def test_ground_score_kernel_matches_torch():
    from dd_b200.functional import ground_score

    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W, max_it = 3, 40, 64, 7
    pts = torch.randn(B, 3, H, W, device="cuda", generator=g)
    w = torch.randn(B * max_it, 3, device="cuda", generator=g)
    row0, tol = 24, 0.3
    got = ground_score(pts, w, row0, tol)
    flat = pts[:, :, row0:, :].reshape(B, 3, -1)
    ref = torch.empty(B * max_it, dtype=torch.int32)
    for k in range(B * max_it):
        x, y, z = flat[k % B]
        ref[k] = int(((x * w[k, 0] + z * w[k, 1] + w[k, 2] - y).abs() < tol).sum())
    assert (got.cpu() - ref).abs().max().item() <= 1   # a threshold comparison may flip in the last ulp


@pytest.mark.parametrize("phase", ["disp_init", "fine_tune"])
def test_ragged_last_batch(phase):
    """A last batch smaller than opt.batch_size must work (the reference slices its per-batch buffers with [:B],
    tools.py:191-197) and give, per image, what the full batch gives: losses are means, so the loss of the first
    image alone must equal the single-image evaluation, and the step on a 1-image batch must run end to end."""
    import options
    from Trainer import Trainer
    from dd_b200 import synthetic

    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "--height", "64",
                                              "--width", "96", "-b", "3", "--g_d_ground", "0.0"])
    opt.cuda_ids, opt.local_rank, opt.ddp = [0], 0, False
    torch.manual_seed(3)
    tr = Trainer(opt)
    tr.setup_phase(phase)
    tr.bool_automask = False          # no device-side tie-break noise: deterministic comparison
    tr.step, tr.num_steps_per_epoch = 100, 100
    tr.set_eval()                      # BatchNorm running statistics, no DropPath: per-image results independent of the batch
    full = {k: v.cuda() for k, v in synthetic.make_batch(opt, 4).items()}
    one = {k: v[:1].contiguous() for k, v in full.items()}
    with torch.no_grad():
        _, l_full = tr.process_batch(dict(full))
        _, l_one = tr.process_batch(dict(one))
        per_image = []
        for i in range(3):
            _, li = tr.process_batch({k: v[i:i + 1].contiguous() for k, v in full.items()})
            per_image.append(float(li["loss_term/p_photo"]))
    assert float(l_one["loss_term/p_photo"]) == pytest.approx(per_image[0], rel=1e-6)
    assert float(l_full["loss_term/p_photo"]) == pytest.approx(sum(per_image) / 3, rel=2e-5)
    tr.set_train()
    outputs, losses = tr.train_step(dict(one))     # backward + optimiser on the ragged batch
    assert torch.isfinite(losses["loss"]).item()
