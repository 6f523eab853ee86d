"""Training orchestration with the reference's Trainer surface (Trainer.py:19-756 in the reference).

Same public methods and dict contracts (`process_batch`, `generate_images_pred`, `compute_losses`,
`setup_phase`, `get_optim`, `run_epoch`, `train`, `save_model`, `load_model`, `vis`-free logging), but the
per-step arithmetic is executed by hand-written sm_100a kernels:

  generate_images_pred  -> ONE launch of dd_warp_photo_fwd for all pyramid levels and both source
                           frames (interp, disp_to_depth, BackprojectDepth, Project3D, scene-flow
                           composition, grid_sample, SSIM+L1, automask min, c_consistency, disp_mag)
  compute_losses        -> assembles the reference's loss dictionary from the kernel's per-level sums,
                           plus one batched dd_smooth_* launch and dd_msparsity_* per (level, frame)
  backward              -> dd_warp_photo_bwd / dd_smooth_bwd / dd_msparsity_bwd / dd_conv_bwd through
                           torch.autograd.Function wrappers; encoders stay in PyTorch/cuDNN.

Data parallelism: one process per GPU; gradients of the phase's parameters live in one flat arena
that is all-reduced in place over NCCL (dd_b200.parallel.GradArena) instead of DistributedDataParallel
buckets.  Data loading / wandb visualisation of the reference are out of scope;
`run_epoch` accepts any iterable of `inputs` dicts.
"""
import json
import os
import os.path as osp
import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

import networks
from dd_b200 import _lib as L
from dd_b200 import checkpoint as ckpt
from dd_b200 import functional as Fn
from dd_b200.parallel import GradArena, broadcast_module_state
from tools import BackprojectDepth, DepthMetrics, GroundPlane, Project3D, SSIM, depth_to_disp, disp_to_depth
from utils import join_dir, sec_to_hm_str

LOSS_PREFIX = "g_"


class Trainer:
    def __init__(self, options):
        self.opt = options
        assert self.opt.height % 32 == 0, f"height(={self.opt.height}) must be a multiple of 32"
        assert self.opt.width % 32 == 0, f"width(={self.opt.width}) must be a multiple of 32"
        assert self.opt.frame_ids[0] == 0, f"frame_ids(={self.opt.frame_ids}) must start with 0"
        assert len(self.opt.epoch_schedules) == 4 and all(e >= 0 for e in self.opt.epoch_schedules)
        assert len(self.opt.frame_ids) - 1 <= L.DD_MAX_FRAMES and len(self.opt.scales) <= L.DD_MAX_SCALES

        self.local_rank = self.opt.local_rank
        self.cuda_id = self.opt.cuda_ids[self.local_rank]
        if not torch.cuda.is_available():
            raise L.DynamoB200Error("Trainer needs a CUDA device: the hot path has no CPU fallback")
        assert self.cuda_id < torch.cuda.device_count(), f"cuda_ids[local_rank](={self.cuda_id}) must be visible"
        self.device = torch.device(f"cuda:{self.cuda_id}")
        torch.cuda.set_device(self.device)
        L.load()

        self.base_model = networks.Model(self.opt)
        if self.opt.load_ckpt != "":
            self.load_model()
        self.base_model.to(self.device)
        self.base_model.prepare_memory_format()     # before any gradient arena exists (see ResnetEncoder.to_channels_last)
        self.model = self.base_model          # no DDP wrapper: gradients are reduced through the arena
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if getattr(self.opt, "ddp", False) else 1
        if self.world_size > 1:
            # What the DDP constructor does in the reference (Trainer.py:44): every replica starts from rank 0's
            # parameters and buffers, whatever the local seeds / initialisation order were.
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("opt.ddp is set with WORLD_SIZE > 1 but torch.distributed is not initialised (train.py does it)")
            broadcast_module_state(self.base_model, src=0)

        self.num_scales = len(self.opt.scales)
        self.B, self.H, self.W = self.opt.batch_size, self.opt.height, self.opt.width
        self.log_path = osp.join(self.opt.log_dir, self.opt.model_name)

        self.depth_metrics = DepthMetrics(self.opt.eval_img_bound, self.opt.eval_min_depth, self.opt.eval_max_depth)
        self.gplane = GroundPlane(num_points_per_it=self.opt.gp_np_per_it, max_it=self.opt.gp_max_it, tol=self.opt.gp_tol,
                                  g_prior=self.opt.gp_prior)
        self.ssim = SSIM().to(self.device)
        self.bce = nn.BCEWithLogitsLoss()
        self.I = torch.eye(4, device=self.device).reshape(1, 4, 4).repeat(self.B, 1, 1)

        import torchvision.transforms as T
        self.resize, self.backproject_depth, self.project_3d = {}, {}, {}
        for s in self.opt.scales:
            h, w = self.H // 2**s, self.W // 2**s
            self.resize[s] = T.Resize((h, w), interpolation=T.InterpolationMode.BICUBIC, antialias=True)
            self.backproject_depth[s] = BackprojectDepth(self.B, h, w).to(self.device)
            self.project_3d[s] = Project3D(self.B, h, w).to(self.device)

        # knobs that do not exist in the reference
        self.materialise_outputs = False     # True: also write the reference's full-resolution by-products into `outputs`
        self.automask_noise = None           # {scale: (B,F,H,W)} to inject the tie-break noise (parity tests)
        self.freeze_inactive = True          # no backward through networks the current phase does not optimise
        self.step, self.epoch, self.g_step = 0, 0, 0
        # steps per epoch drive the loss-weight ramp (Trainer.py:307-308); the reference takes len(train_loader)
        # (Trainer.py:505), run_epoch does the same whenever the loader has a length
        self.num_steps_per_epoch = max(1, int(getattr(self.opt, "epoch_size", 1)))
        self.bool_automask = False
        self.arena = None
        self._arena_checked = False
        self.optim = None
        # all-reduce arena chunks from autograd hooks while backward is still running (DD_OVERLAP=0: after backward)
        self.overlap_allreduce = os.environ.get("DD_OVERLAP", "1") != "0"
        if self.is_main() and getattr(self.opt, "model_name", "--") != "--":
            self.save_opt()                  # as the reference (Trainer.py:85)

    # ------------------------------------------------------------------ phases / optimiser
    def setup_phase(self, phase_name):
        table = {"disp_init": (False, False, ["Depth", "Pose"], 1.0),
                 "motion_init": (True, False, ["CmpFlow"], 1.0),
                 "mask_init": (True, True, ["Pose", "CmpFlow", "MotMask"], 1.0),
                 "fine_tune": (True, True, ["Depth", "Pose", "CmpFlow", "MotMask"], 0.5)}
        if phase_name not in table:
            raise Exception(f"Phase name {phase_name} not recognized.")
        cmp_flow, mot_mask, nets, lr_factor = table[phase_name]
        self.base_model.bool_CmpFlow, self.base_model.bool_MotMask = cmp_flow, mot_mask
        if self.freeze_inactive:
            # Only the phase's networks are stepped (Trainer.py:466-490); the reference still back-propagates
            # into all of them.  Skipping those gradients changes no optimised weight.
            active = set(m for n in nets for m in self.base_model.network2modules[n])
            for name in self.base_model.module_names:
                for p in getattr(self.base_model, name).parameters():
                    p.requires_grad_(name in active)
                    if name not in active:
                        p.grad = None
        self.optim = self.get_optim(nets, lr_factor=lr_factor)
        self.phase_name = phase_name

    def get_optim(self, network_names, optm=optim.Adam, lr_factor=1):
        named = self.base_model.named_parameters_by_names(network_names)
        params = [p for _, p in named]
        if self.arena is not None:
            self.arena.release()
        self.arena = GradArena(params, world_size=self.world_size, names=[n for n, _ in named],
                               chunk_ids=[n.split(".", 1)[0] for n, _ in named], overlap=self.overlap_allreduce)
        self.param_names = self.arena.names
        self._arena_checked = False
        kw = {"fused": True} if optm is optim.Adam else {}
        optimizer = optm(params, self.opt.learning_rate * lr_factor, **kw)
        sched = optim.lr_scheduler.StepLR(optimizer, self.opt.scheduler_step_size, 0.5)
        return {"optimizer": optimizer, "lr_scheduler": sched, "network_names": network_names}

    # ------------------------------------------------------------------ training loop
    def train(self, loader_factory=None, resume_from=None):
        """loader_factory(trainer) -> iterable of `inputs` dicts for one epoch (the reference's dataset
        classes are out of scope; dd_b200.synthetic.SyntheticTriplets is the built-in source).
        resume_from (or options --resume): a `models/<phase>_<epoch>` folder written by save_model; the run continues
        with the next epoch of that phase (weights, Adam moments, LR schedule, counters and RNG streams restored)."""
        if loader_factory is None:
            from dd_b200.synthetic import SyntheticTriplets
            loader_factory = lambda tr: SyntheticTriplets(tr.opt, steps=tr.num_steps_per_epoch, device=tr.device)
        self.loader_factory = loader_factory
        self.g_step = 0
        resume_from = resume_from or getattr(self.opt, "resume", "")
        first_phase, first_epoch, state = 0, 0, None
        if resume_from:
            state = ckpt.load_state(resume_from)
            first_phase, first_epoch = ckpt.resume_point(state, self.opt.epoch_schedules)
            self.opt.load_ckpt = resume_from
            self.load_model()
            self.base_model.to(self.device)
            self.g_step = state["g_step"]
            self.print(f"======== resuming after {state['phase_name']} epoch {state['epoch']} ({resume_from}) ========")
        if state is not None:
            ckpt.restore_rng_state(state)          # also when the checkpoint closed a phase and the next one starts fresh
        for phase_i, phase_name in enumerate(ckpt.PHASES):
            num_epoch = self.opt.epoch_schedules[phase_i]
            if phase_i < first_phase:
                continue
            self.print(f"======== {phase_name.upper()} - Num Epochs={num_epoch} ========")
            if num_epoch > 0:
                same_phase = state is not None and phase_name == state["phase_name"]
                self.run_phase(phase_name, num_epoch, start_epoch=first_epoch if phase_i == first_phase else 0,
                               state=state if same_phase else None)

    def run_phase(self, phase_name, num_epoch, start_epoch=0, state=None):
        self.setup_phase(phase_name)
        self.step, self.epoch = 0, 0
        if state is not None:          # continuing inside the phase the checkpoint was written in
            ckpt.apply_state(state, self.optim["optimizer"], self.optim["lr_scheduler"], param_names=self.param_names,
                             restore_rng=False)
            self.step = state["step"]
        self.bool_automask = phase_name == "disp_init"
        self.start_time = time.time()
        for self.epoch in range(start_epoch, num_epoch):
            loader = self.loader_factory(self)
            if hasattr(loader, "__len__"):
                if len(loader) <= 0:
                    raise ValueError("the loader of this epoch is empty")
                self.num_steps_per_epoch = len(loader)     # Trainer.py:505
            self.num_total_steps = self.num_steps_per_epoch * num_epoch
            self.run_epoch(loader)
            if ((self.epoch + 1) % self.opt.save_frequency == 0) or (self.epoch == num_epoch - 1):
                self.save_model(phase_name)

    def run_epoch(self, loader):
        self.set_train()
        self.optim["optimizer"].zero_grad(set_to_none=False)
        t0 = time.time()
        for batch_idx, inputs in enumerate(loader):
            outputs, losses = self.train_step(inputs)
            if batch_idx % self.opt.log_frequency == 0 and self.is_main():
                dt = time.time() - t0
                self.print(f"epoch {self.epoch:>3} | batch {batch_idx:>6} | loss {float(losses['loss'].detach()):.5f} | "
                           f"{sec_to_hm_str(dt)}")
            del outputs
        self.optim["lr_scheduler"].step()

    def train_step(self, inputs):
        """process_batch + backward + gradient all-reduce + Adam (reference: Trainer.py:147-151)."""
        outputs, losses = self.process_batch(inputs)
        losses["loss"].backward()
        if not self._arena_checked:       # once per phase: every .grad must still alias the arena, or the exchange is void
            if not self.arena.check_views():
                raise RuntimeError("a parameter's .grad no longer aliases the gradient arena (was a module converted with "
                                   ".to(memory_format=...) / .to(dtype) after setup_phase?)")
            self._arena_checked = True
        self.arena.all_reduce()
        self.optim["optimizer"].step()
        self.optim["optimizer"].zero_grad(set_to_none=False)
        self.g_step += 1
        self.step += 1
        return outputs, losses

    def process_batch(self, inputs):
        self.process_inputs(inputs)
        outputs = self.model(inputs)
        self.generate_images_pred(inputs, outputs)
        losses = self.compute_losses(inputs, outputs)
        return outputs, losses

    # ------------------------------------------------------------------ fused view synthesis
    def generate_images_pred(self, inputs, outputs):
        opt, bm = self.opt, self.base_model
        frames = opt.frame_ids[1:]
        scales = list(opt.scales)
        want = ("warped", "sample", "depth", "ident_sel", "resid", "independ") if self.materialise_outputs else ()
        cfg = Fn.WarpConfig(scales=scales, cmpflow=bm.bool_CmpFlow, motmask=bm.bool_MotMask, automask=self.bool_automask,
                            min_depth=opt.min_depth, max_depth=opt.max_depth, ssim_weight=opt.ssim_weight,
                            mask_disp_thrd=opt.mask_disp_thrd, materialise=want)
        target = inputs[("color", 0, 0)]
        B = target.shape[0]
        disps = [outputs[("disp", 0, s)] for s in scales]
        flows = [[outputs[("complete_flow", f, s)] for f in frames] for s in scales] if bm.bool_CmpFlow else None
        masks = [[outputs[("motion_mask", f, s)] for f in frames] for s in scales] if bm.bool_MotMask else None
        noises = None
        if self.bool_automask:   # tie-break noise of Trainer.py:339, drawn on the device
            if self.automask_noise is not None:
                noises = [self.automask_noise[s].to(self.device) for s in scales]
            else:
                noises = [torch.randn(B, len(frames), self.H, self.W, device=self.device) for _ in scales]
        sums = Fn.view_synthesis_sums(cfg, target, [inputs[("color", f, 0)] for f in frames], inputs[("K", 0)],
                                      inputs[("inv_K", 0)], [outputs[("cam_T_cam", 0, f)] for f in frames],
                                      [inputs[("ts", f)] for f in frames], disps, flows, masks, noises)
        outputs["_dd_sums"] = sums
        outputs["_dd_aux"] = cfg.aux
        if not bm.bool_MotMask:
            for s in scales:
                for f in frames:   # Trainer.py:244-245: constant-one mask when no mask network is active
                    h, w = disps[scales.index(s)].shape[-2:]
                    outputs[("motion_mask", f, s)] = torch.ones(B, 1, h, w, device=self.device)
        if self.materialise_outputs:
            self._publish_aux(inputs, outputs, cfg, frames, scales)

    def _publish_aux(self, inputs, outputs, cfg, frames, scales):
        """Fill `outputs` with the by-products the reference materialises (Trainer.py:229-287)."""
        for (name, fi, li), t in cfg.aux.items():
            s = scales[li]
            if name == "depth":
                outputs[("depth", 0, s)] = t
                outputs[("disp_scaled", 0, s)] = 1 / t
            elif name == "ident_sel":
                outputs[f"identity_selection/{s}"] = t
            else:
                key = {"warped": "color", "sample": "sample", "resid": "residual_flow", "independ": "independ_flow"}.get(name)
                if key:
                    outputs[(key, frames[fi], s)] = t
        if self.bool_automask:
            for s in scales:
                for f in frames:
                    outputs[("color_identity", f, s)] = inputs[("color", f, 0)]

    # ------------------------------------------------------------------ loss assembly
    def loss_coefficients(self):
        """loss_term name -> coefficient with the linear ramp of Trainer.py:299-310."""
        coefs = {}
        for k, v in vars(self.opt).items():
            if k.startswith(LOSS_PREFIX):
                val = v
                if k in self.opt.weight_ramp:
                    val = val * float(np.clip(self.opt.ramp_red * self.step / self.num_steps_per_epoch, 0.0, 1.0))
                coefs[k[len(LOSS_PREFIX):]] = val
        return coefs

    def compute_losses(self, inputs, outputs):
        opt, bm = self.opt, self.base_model
        nets = self.optim["network_names"]
        move_Depth, move_CmpFlow, move_MotMask = "Depth" in nets, "CmpFlow" in nets, "MotMask" in nets
        frames, scales = opt.frame_ids[1:], list(opt.scales)
        nf = len(frames)
        coef = self.loss_coefficients()
        terms = list(coef.keys())
        losses = {"loss": 0}
        for t in terms + scales:
            losses[f"loss_term/{t}"] = 0
        for t in terms:
            losses[f"loss_coef/{t}"] = coef[t]

        sums, aux = outputs["_dd_sums"], outputs["_dd_aux"]
        B, H, W = inputs[("color", 0, 0)].shape[0], self.H, self.W

        # one batched smoothness launch for every (term, level, frame)
        tasks, task_keys = [], []
        for li, s in enumerate(scales):
            color = inputs[("color", 0, s)]
            if move_Depth and coef.get("d_smooth", 0) > 0:
                tasks.append((outputs[("disp", 0, s)], color, True)); task_keys.append(("d_smooth", li))
            for f in frames:
                if move_CmpFlow and bm.bool_CmpFlow and coef.get("c_smooth", 0) > 0:
                    tasks.append((outputs[("complete_flow", f, s)], color, False)); task_keys.append(("c_smooth", li))
                if move_MotMask and bm.bool_MotMask and coef.get("m_smooth", 0) > 0:
                    tasks.append((outputs[("motion_mask", f, s)], color, False)); task_keys.append(("m_smooth", li))
        smooth_vals = None
        if tasks:
            ssums = Fn.smooth_sums([t[0] for t in tasks], [t[1] for t in tasks], [t[2] for t in tasks])
            smooth_vals = Fn.smooth_means(ssums, [tuple(t[0].shape) for t in tasks])

        for li, s in enumerate(scales):
            h, w = H >> s, W >> s
            ps = {t: 0 for t in terms}
            ps["p_photo"] = sums[li, L.DD_SUM_PHOTO] / (B * H * W)
            if self.bool_automask and ("ident_sel", 0, li) in aux:
                outputs[f"identity_selection/{s}"] = aux[("ident_sel", 0, li)]
            if smooth_vals is not None:
                for ti, (name, lj) in enumerate(task_keys):
                    if lj == li:
                        ps[name] = ps[name] + smooth_vals[ti] / (2**s) / (1 if name == "d_smooth" else nf)
            if move_Depth and coef.get("d_ground", 0) > 0 and bm.bool_MotMask:
                _, disp_diff, _ = self.process_ground(inputs, outputs, scale=s)
                disp_diff = torch.where(disp_diff > 0, torch.zeros_like(disp_diff), disp_diff)
                ps["d_ground"] = -1 * torch.mean(disp_diff) / (2**s)
            for fi, f in enumerate(frames):
                if move_CmpFlow and bm.bool_CmpFlow and bm.bool_MotMask and coef.get("c_consistency", 0) > 0:
                    ps["c_consistency"] = ps["c_consistency"] + sums[li, L.DD_SUM_CONSIST0 + fi] / (B * 3 * h * w) / (2**s) / nf
                if move_MotMask and bm.bool_MotMask and coef.get("m_sparsity", 0) > 0:
                    mag_sum = sums[li, L.DD_SUM_MAG0 + fi].detach().reshape(1)
                    sp = Fn.motion_sparsity(aux[("mag", fi, li)], mag_sum, outputs[("motion_prob", f, s)])
                    ps["m_sparsity"] = ps["m_sparsity"] + sp / (2**s) / nf
            for t in terms:
                losses[f"loss_term/{s}"] = losses[f"loss_term/{s}"] + ps[t] * coef[t]
                losses[f"loss_term/{t}"] = losses[f"loss_term/{t}"] + ps[t]
            losses["loss"] = losses["loss"] + losses[f"loss_term/{s}"] / self.num_scales
        return losses

    def compute_reprojection_loss(self, pred, target):
        """Stand-alone per-pixel reprojection loss (reference: Trainer.py:413-423) through the SSIM kernel."""
        l1 = torch.abs(target - pred).mean(1, True)
        return self.opt.ssim_weight * self.ssim(pred, target).mean(1, True) + (1 - self.opt.ssim_weight) * l1

    # ------------------------------------------------------------------ ground prior (host driven)
    def process_ground(self, inputs, outputs, scale=0):
        disp = outputs[("disp", 0, scale)]
        _, depth = disp_to_depth(disp, self.opt.min_depth, self.opt.max_depth)
        inv_K = inputs[("inv_K", scale)]
        h, w = self.H // 2**scale, self.W // 2**scale
        cam = self.backproject_depth[scale](depth, inv_K)
        plane_dist, plane_param = self.gplane(cam[:, :3].reshape(-1, 3, h, w))
        g_mask = (plane_dist.abs() < self.opt.gp_tol).float()
        shifted = plane_param.clone()
        shifted[:, 2] += self.opt.gp_tol
        ground_disp, ground_depth = self.get_ground_depth(shifted, inv_K, scale)
        disp_diff = disp - ground_disp
        disp_diff = torch.where(ground_depth == self.opt.max_depth, torch.zeros_like(disp_diff), disp_diff)
        return plane_dist, disp_diff, g_mask

    def get_ground_depth(self, plane_param, inv_K, scale=0):
        h, w = self.H // 2**scale, self.W // 2**scale
        B = inv_K.size(0)
        ones = torch.ones(B, 1, h, w, device=inv_K.device)
        rays = self.backproject_depth[scale](ones, inv_K)[:, :3]          # inv_K[:3,:3] @ (u,v,1)
        w1, w2, w3 = plane_param[:, 0:1], plane_param[:, 1:2], plane_param[:, 2:3]
        vx, vy, vz = rays[:, 0:1], rays[:, 1:2], rays[:, 2:3]
        gd = (w3 / (vy - vx * w1 - vz * w2)).reshape(B, 1, h, w)
        gd = torch.where((gd < 0) | (gd > self.opt.max_depth), torch.full_like(gd, self.opt.max_depth), gd)
        return depth_to_disp(gd, self.opt.min_depth, self.opt.max_depth), gd

    # ------------------------------------------------------------------ inputs / checkpoints / misc
    def prefetch(self, inputs):
        """Start the host -> device copies of a (pinned) batch on a dedicated copy stream so that they overlap the
        training step in flight; returns the device-side dict to hand to train_step / process_batch, which waits for
        the copies (an event, no host synchronisation) before touching the tensors."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        dev, moved = {}, {}
        with torch.cuda.stream(self._copy_stream):
            for key, inp in inputs.items():
                if torch.is_tensor(inp) and inp.device != self.device:
                    if id(inp) not in moved:
                        moved[id(inp)] = inp.to(self.device, non_blocking=True)
                    dev[key] = moved[id(inp)]
                else:
                    dev[key] = inp
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        dev["__ready__"] = (ready, list(moved.values()))
        return dev

    def process_inputs(self, inputs):
        """host -> device first (non-blocking from pinned memory), then the colour pyramid on the GPU; the
        reference resizes on the CPU before the copy (Trainer.py:722-727), which sits on the critical path."""
        pending = inputs.pop("__ready__", None)
        if pending is not None:      # batch staged by prefetch(): order the compute stream after the copies
            ready, tensors = pending
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ready)
            for t in tensors:
                t.record_stream(cur)
        moved = {}
        for key, inp in inputs.items():
            if torch.is_tensor(inp) and inp.device != self.device:
                if id(inp) not in moved:      # color_aug aliases color when no augmentation was drawn
                    moved[id(inp)] = inp.to(self.device, non_blocking=True)
                inputs[key] = moved[id(inp)]
        self.apply_img_resize(inputs)

    def apply_img_resize(self, inputs):
        """('color',0,s) for s>0: chained bicubic-antialias x1/2, clamped (reference: Trainer.py:729-734)."""
        for s in self.opt.scales:
            if s != 0 and ("color", 0, s) not in inputs:
                prev = inputs[("color", 0, s - 1)]
                if prev.is_cuda and prev.shape[-2] % 2 == 0 and prev.shape[-1] % 2 == 0 and \
                        tuple(prev.shape[-2:]) == (2 * (self.H >> s), 2 * (self.W >> s)):
                    inputs[("color", 0, s)] = Fn.pyramid_half(prev)       # hand-written kernel (dd_pyramid_half_fwd)
                else:                                                     # host-side tensors (data-loader workers)
                    inputs[("color", 0, s)] = torch.clamp(self.resize[s](prev), 0, 1)

    def save_opt(self):
        folder = join_dir(self.log_path, "models")
        with open(osp.join(folder, "opt.json"), "w") as fh:
            json.dump({k: v for k, v in vars(self.opt).items()}, fh, indent=2, default=str)

    def save_model(self, save_name):
        if not self.is_main():
            return
        folder = join_dir(self.log_path, "models", f"{save_name}_{self.epoch:02}")
        self.base_model.save(folder)
        torch.save(self.optim["optimizer"].state_dict(), osp.join(folder, "adam.pth"))   # as the reference (Trainer.py:706-707)
        ckpt.save_state(folder, ckpt.pack_state(self.phase_name, self.epoch, self.step, self.g_step, self.optim["optimizer"],
                                                self.optim["lr_scheduler"], self.opt.epoch_schedules,
                                                param_names=[(n, tuple(p.shape)) for n, p in zip(self.arena.names, self.arena.params)]))

    def load_model(self):
        self.base_model.load(verbose=self.is_main())

    def is_main(self):
        return self.local_rank == 0

    def print(self, s=""):
        if self.is_main():
            print(s)

    def set_train(self):
        self.base_model.set_train()

    def set_eval(self):
        self.base_model.set_eval()
