"""Benchmark of the per-step hot path (BASELINE.json): frame-triplets/s of one training step, plus the HBM roofline
of the fused warp+SSIM kernel, the same step of the UNMODIFIED reference in eager PyTorch on the same B200 ("the real
bar") and the reference's CPU path timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--phase P] [--impl ours|reference|eager]

  --config litemono_bs32      (default, BASELINE configs[1]/[3]) litemono Waymo-shape 192x640 bs32 per GPU
           md2_bs16           (configs[2]) monodepthv2 KITTI-shape 192x640 bs16, 4-scale loss
           lite_384x768_bs16  (configs[4]) litemono nuScenes-shape 384x768 bs16 per GPU (float time steps, scene flow)
  --phase  fine_tune (default: all four networks, every loss term) | disp_init | motion_init | mask_init
  --impl   ours       the product (hand-written sm_100a kernels behind the reference's Trainer surface)
           reference  the reference's CPU implementation on the host cores (the unmodified reference from
                      baseline/_ref when build() installed it, else the oracle port), bounded sample per step
           eager      the unmodified reference in eager PyTorch on cuda:0 (helper of `variants.eager_b200`)

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dynamo-depth_b200")
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
if "reference" in [a for i, a in enumerate(sys.argv) if i > 0 and sys.argv[i - 1] == "--impl"]:
    os.environ["CUDA_VISIBLE_DEVICES"] = ""          # the CPU arm must not see the GPU (the reference picks cuda when it can)
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "frame_triplets_per_sec_training_step"
UNIT = "triplets/s"

CONFIGS = {
    "litemono_bs32": dict(model="litemono", kind="waymo", H=192, W=640, batch=32,
                          label="litemono Waymo-shape 192x640 bs32 synthetic"),
    "md2_bs16": dict(model="monodepthv2", kind="kitti", H=192, W=640, batch=16,
                     label="monodepthv2 KITTI-shape 192x640 bs16 synthetic, 4-scale loss"),
    "lite_384x768_bs16": dict(model="litemono", kind="nuscenes", H=384, W=768, batch=16,
                              label="litemono nuScenes-shape 384x768 bs16 synthetic + motion_decoder scene-flow"),
}


def argv_for(cfg, batch):
    return ["-d", cfg["kind"], "--depth_model", cfg["model"], "--weights_init", "scratch", "--height", str(cfg["H"]),
            "--width", str(cfg["W"]), "-b", str(batch)]


def make_opt(cfg, batch, local_rank=0, extra=()):
    import options

    opt = options.DynamoOptions().parse(args=argv_for(cfg, batch) + list(extra))
    opt.ddp = int(os.environ.get("WORLD_SIZE", 1)) > 1
    opt.local_rank = 0
    opt.cuda_ids = [local_rank]
    return opt


def workload_name(cfg, phase):
    return f"{cfg['label']}, phase {phase}"


def algorithmic_bytes(cfg, batch, scales, flow_mask):
    """per image and level fwd = 36*P + 4*P/4^s (+16*P/4^s with flow+mask);
    bwd = fwd + 4*P/4^s (+32*P/4^s instead of +16 with flow+mask)."""
    P = cfg["H"] * cfg["W"]
    fwd = sum(36 * P + 4 * P / 4**s + (16 * P / 4**s if flow_mask else 0) for s in scales)
    bwd = sum(36 * P + 8 * P / 4**s + (32 * P / 4**s if flow_mask else 0) for s in scales)
    return batch * fwd, batch * bwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arms
def reference_root():
    """Where the unmodified reference lives on THIS box: baseline/_ref (installed by __graft_entry__.build() in the
    build container, git-ignored, travels with gpurun) -- never /root/reference, which does not exist on the GPU box."""
    root = os.environ.get("DD_REFERENCE_ROOT", REF_INSTALL)
    return root if os.path.isfile(os.path.join(root, "Trainer.py")) else None


def reference_trainer(cfg, batch, phase, device_kind):
    """(step_fn(inputs) -> loss tensor, opt, kind): one optimisation step of the reference (Trainer.py:147-151) --
    the unmodified reference through oracle/_refshim.py when it is installed, else the oracle port."""
    from dd_b200 import synthetic  # noqa: F401  (imported before the reference's module names shadow the product's)

    root = reference_root()
    if root is not None:
        os.environ["DD_REFERENCE_ROOT"] = root
        from oracle import _refshim
        ns = _refshim.load_reference()
        tr = _refshim.make_reference_trainer(ns, argv_for(cfg, batch), phase=phase, step=100, steps_per_epoch=100)
        assert str(tr.device).startswith(device_kind), (tr.device, device_kind)
        tr.set_train()

        def step(inputs):
            outputs, losses = tr.process_batch(inputs)
            losses["loss"].backward()
            tr.optim["optimizer"].step()
            tr.optim["optimizer"].zero_grad()
            del outputs
            return losses["loss"]
        return step, tr.opt, "reference"
    # oracle port (plain torch restatement of the same step) around the product's PyTorch encoders
    import networks
    from oracle import networks as on
    from oracle.synth import add_color_pyramid

    opt = make_opt(cfg, batch, extra=["--encoder_linear", "torch"])
    prod = networks.Model(opt)
    states = {n: {k: v for k, v in getattr(prod, n).state_dict().items() if not k.startswith("net.")}
              for n in ("depth_dec", "pose_dec", "motion_dec", "motion_mask")}
    om = on.OracleModel(cfg["model"], opt.scales, opt.frame_ids, prod.depth_enc, prod.pose_enc, prod.motion_enc, states)
    dev = torch.device("cuda:0" if device_kind == "cuda" else "cpu")
    om.to(dev)
    om.train()
    tr = on.OracleTrainer(om, cfg["H"], cfg["W"], learning_rate=opt.learning_rate)
    tr.setup_phase(phase)
    tr.step, tr.steps_per_epoch = 100, 100

    def step(inputs):
        inputs = {k: v.to(dev) for k, v in inputs.items()}
        add_color_pyramid(inputs, opt.scales, cfg["H"], cfg["W"])
        return tr.train_step(inputs)
    return step, opt, "port"


def run_reference(args):
    """The reference's own CPU implementation of the step on the host cores, a bounded sample per step."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cfg = CONFIGS[args.config]
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1234)
    sample_b = args.sample_batch or (2 if (args.steps + args.warmup) <= 12 and cfg["H"] * cfg["W"] <= 192 * 640 else 1)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):        # the reference prints banners; stdout carries the JSON line only
        step, opt, kind = reference_trainer(cfg, sample_b, args.phase, "cpu")
        from dd_b200 import synthetic
        batch = synthetic.make_batch(opt, 1234, batch=sample_b, kind=cfg["kind"])
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            step(dict(batch))
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total, cores = sum(times), torch.get_num_threads()
    value = sample_b * len(times) / total
    what = ("unmodified reference (baseline/_ref) Trainer.process_batch + backward + Adam" if kind == "reference"
            else "oracle port of the reference step")
    sample = (f"{len(times)} timed steps of {sample_b} triplets each after {args.warmup} warm-up, {what}, all loss terms of phase "
              f"{args.phase}; bounded sample of the bs{cfg['batch']} step")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.phase), "sample_batch": sample_b},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_eager(args):
    """The unmodified reference, eager PyTorch, on cuda:0, same batch shape / phase / loss terms, inputs resident on the
    device.  `--tf32 0` switches cuDNN / cuBLAS TF32 off (fp32 arithmetic as the parity tests use), `--tf32 1` is torch's
    default (cuDNN convolutions in TF32)."""
    cfg = CONFIGS[args.config]
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    if not args.tf32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        step, opt, kind = reference_trainer(cfg, cfg["batch"], args.phase, "cuda")
        from dd_b200 import synthetic
        batches = [{k: v.cuda() for k, v in synthetic.make_batch(opt, 1234 + i, batch=cfg["batch"], kind=cfg["kind"]).items()}
                   for i in range(2)]
        for i in range(max(args.warmup, 3)):
            step(dict(batches[i % 2]))
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(args.steps):
            loss = step(dict(batches[i % 2]))
        end.record()
        torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    print(json.dumps({"impl": "eager_b200", "kind": kind, "tf32": bool(args.tf32), "value": cfg["batch"] * args.steps / (ms / 1000),
                      "unit": UNIT, "ms_per_step": ms / args.steps, "steps": args.steps, "warmup": max(args.warmup, 3),
                      "loss": float(loss.detach()), "workload": workload_name(cfg, args.phase),
                      "note": ("unmodified reference (baseline/_ref), eager PyTorch %s on the same B200, inputs resident on the device"
                               % torch.__version__) if kind == "reference" else "oracle port (plain torch) on the same B200"}), flush=True)


def _child_json(argv, timeout, env=None):
    """Run `python bench.py ...` in a child process (the reference's module names collide with the product's) and parse
    the JSON line it prints; a failure is reported, never raised."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, capture_output=True, text=True, timeout=timeout,
                             env=env or os.environ.copy())
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------ the product
def run_ours(args):
    import torch.distributed as dist
    from Trainer import Trainer
    from dd_b200 import _lib as L
    from dd_b200 import functional as Fn
    from dd_b200 import synthetic

    cfg = CONFIGS[args.config]
    BATCH = cfg["batch"]
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.backends.cudnn.benchmark = True
    lib = L.load()

    opt = make_opt(cfg, BATCH, local_rank)
    torch.manual_seed(1234)                # identical weights on every rank (Trainer also broadcasts rank 0's); data: + rank
    tr = Trainer(opt)
    tr.setup_phase(args.phase)
    tr.bool_automask = args.phase == "disp_init"
    tr.num_steps_per_epoch = 100
    tr.step = 100                      # loss-weight ramp = 1 (options.py:106-114)
    tr.set_train()

    dev = tr.device
    resident = synthetic.SyntheticTriplets(opt, steps=1, device=dev, seed=1234 + 17 * rank, distinct=2).batches
    pinned = synthetic.SyntheticTriplets(opt, steps=1, device=None, seed=1234 + 17 * rank, pinned=True, distinct=2).batches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, e2e):
        barrier()
        launches0 = lib.dd_launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        last = None
        staged = tr.prefetch(dict(batches[0])) if e2e else None
        for i in range(steps):
            if e2e:   # this step's batch was staged by the copy stream; the next one is issued before the step runs
                cur_batch, staged = staged, (tr.prefetch(dict(batches[(i + 1) % len(batches)])) if i + 1 < steps else None)
            else:
                cur_batch = dict(batches[i % len(batches)])
            outputs, losses = tr.train_step(cur_batch)
            if e2e:
                last = float(losses["loss"].detach().cpu())      # device -> host read of the step's result
            del outputs
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.dd_launch_count() - launches0, last

    # warm-up (cuDNN autotuning, allocator growth), then the device-resident timed region
    timed(resident, max(args.warmup, 3), False)
    if args.profile_step:   # under `ncu --profile-from-start off`: capture exactly one training step
        torch.cuda.profiler.start()
        timed(resident, 1, False)
        torch.cuda.profiler.stop()
        return
    if args.trace_step:     # torch.profiler timeline of two steps: where the GPU waits for the host (dev/gap_report.py reads it)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            timed(resident, 2, False)
        prof.export_chrome_trace(args.trace_step)
        return
    Fn.KERNEL_TIMERS = {}
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, launches, _ = timed(resident, args.steps, False)
    clock_info = clocks.stop() if rank == 0 else None
    kt = Fn.KERNEL_TIMERS
    Fn.KERNEL_TIMERS = None
    k_fwd = statistics.mean(a.elapsed_time(b) for a, b in kt.get("warp_photo_fwd", [])) if kt.get("warp_photo_fwd") else None
    k_bwd = statistics.mean(a.elapsed_time(b) for a, b in kt.get("warp_photo_bwd", [])) if kt.get("warp_photo_bwd") else None

    # gradient exchange (N > 1): device time of the all-reduce tail that is NOT hidden behind backward
    collective = None
    if world > 1:
        tails = []
        for i in range(3):
            outputs, losses = tr.process_batch(dict(resident[i % 2]))
            losses["loss"].backward()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tr.arena.all_reduce()
            b.record()
            tr.optim["optimizer"].step()
            tr.optim["optimizer"].zero_grad(set_to_none=False)
            torch.cuda.synchronize()
            tails.append(a.elapsed_time(b))
            del outputs
        t = torch.tensor([statistics.median(tails)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        collective = {"arena_bytes": tr.arena.numel * 4, "chunks": len(tr.arena.chunks), "collectives_per_step": tr.arena.last_collectives,
                      "op": "ncclAllReduce AVG (1/world inside the collective), chunks issued from autograd hooks during backward",
                      "exposed_tail_ms": float(t.item())}

    # tensor-core linear kernel (csrc/linear_tc.cu): two extra, untimed-for-the-headline steps with per-call events
    lin = {}
    if cfg["model"] == "litemono":
        Fn.KERNEL_TIMERS = {"__detail__": True}
        timed(resident, 2, False)
        kt_lin = Fn.KERNEL_TIMERS
        Fn.KERNEL_TIMERS = None
        for key in ("linear_fwd", "linear_bwd"):
            calls = kt_lin.get(key, [])
            if calls:
                ms_k = sum(a.elapsed_time(b) for a, b, _ in calls)
                lin[key] = {"calls_per_step": len(calls) // 2, "ms_per_step": ms_k / 2,
                            "tflops_fp32_equivalent": sum(m[0] for _, _, m in calls) / (ms_k * 1e-3) / 1e12,
                            "algorithmic_gbs": sum(m[1] for _, _, m in calls) / (ms_k * 1e-3) / 1e9}

    # end-to-end: pinned host inputs copied every step + loss read back every step
    timed(pinned, 4, True)      # the copy stream's staging buffers reach their steady state (three batches in flight)
    ms_e2e, _, _ = timed(pinned, args.steps, True)
    h2d = sum(v.numel() * v.element_size() for v in {id(v): v for v in pinned[0].values()}.values())

    # ---- variants (not the headline).  Every rank runs them (they contain the gradient exchange).
    variants = None
    if not args.no_variants:
        from networks.depth_encoder import EncoderLinear
        v_steps = max(3, args.steps // 2)
        variants = {}

        def measure(name, note, setup, restore):
            setup()
            try:
                timed(resident, 2, False)
                ms_v, _, _ = timed(resident, v_steps, False)
                variants[name] = {"value": world * BATCH * v_steps / (ms_v / 1000), "unit": UNIT, "ms_per_step": ms_v / v_steps,
                                  "steps": v_steps, "note": note}
            finally:
                restore()

        def set_tf32(on):
            torch.backends.cudnn.allow_tf32 = on
            torch.backends.cuda.matmul.allow_tf32 = False     # torch default

        measure("tf32_off", "torch.backends.cudnn.allow_tf32 = False: the encoders' cuDNN convolutions in fp32 as well -- the arithmetic the "
                "parity tests run (the headline keeps torch's default, TF32 cuDNN convolutions, as the reference itself does on a GPU)",
                lambda: set_tf32(False), lambda: set_tf32(True))

        def set_skip(on):
            tr.base_model.skip_unused_depth = on
        measure("skip_unused_depth", "options --skip_unused_depth (SURVEY 8f-3): no depth passes on frames -1/+1, whose disparities no loss term "
                "reads (networks/model.py:69-74 of the reference); optimised-weight gradients unchanged, BatchNorm running statistics differ",
                lambda: set_skip(True), lambda: set_skip(False))

        def set_pose_batch(on):
            tr.base_model.batch_pose_pairs = on
        measure("batch_pose_pairs", "options --batch_pose_pairs (SURVEY 8f-3): both pose-encoder calls (networks/model.py:82-86) as one batch of 2B "
                "pairs; changes the BatchNorm batch statistics of the pose encoder, so opt-in",
                lambda: set_pose_batch(True), lambda: set_pose_batch(False))
        if cfg["model"] == "litemono":
            def set_lin(tf32, mode):
                EncoderLinear.tf32, EncoderLinear.mode = tf32, mode
            measure("encoder_linear_torch_fp32", "options --encoder_linear torch: the Lite-Mono encoder's nn.Linear contractions through torch's "
                    "fp32 matmul (cuBLAS SIMT) instead of csrc/linear_tc.cu", lambda: set_lin(False, "torch"), lambda: set_lin(False, "tc3x"))
            measure("encoder_linear_tf32", "options --encoder_tf32_linear: the same contractions in single-pass TF32 (cuBLAS, reduced "
                    "precision; not parity-grade)", lambda: set_lin(True, "tc3x"), lambda: set_lin(False, "tc3x"))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * BATCH * args.steps / (ms / 1000)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1000)
    flow_mask = args.phase in ("mask_init", "fine_tune")
    fwd_bytes, bwd_bytes = algorithmic_bytes(cfg, BATCH, opt.scales, flow_mask)
    peaks = {}
    peak_src = "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tkey = "" if args.config == "litemono_bs32" else f"_{args.config}"
    roofline = {"bound": "hbm", "kernel": "warp_photo_fwd_kernel (dd_warp_photo_fwd: view synthesis + SSIM/L1 + min, all levels)",
                "achieved": (fwd_bytes / (k_fwd * 1e-3) / 1e9) if k_fwd else None, "peak": peak, "unit": "GB/s",
                "frac": (fwd_bytes / (k_fwd * 1e-3) / 1e9 / peak) if k_fwd else None,
                "traffic": traffic.get(f"warp_photo_fwd_{args.phase}{tkey}"), "algorithmic_bytes": fwd_bytes, "kernel_ms": k_fwd,
                "peak_source": peak_src,
                "backward": {"kernel": "warp_photo_bwd_kernel", "algorithmic_bytes": bwd_bytes, "kernel_ms": k_bwd,
                             "achieved": (bwd_bytes / (k_bwd * 1e-3) / 1e9) if k_bwd else None,
                             "frac": (bwd_bytes / (k_bwd * 1e-3) / 1e9 / peak) if k_bwd else None,
                             "traffic": traffic.get(f"warp_photo_bwd_{args.phase}{tkey}")}}
    if lin:
        # the tensor-core kernel of the step (Lite-Mono linear layers, tcgen05 3xTF32): skinny GEMMs, so HBM is the bound;
        # the tensor figure is fp32-equivalent work (each counted multiply-add costs three TF32 MMAs) against bf16_tflops / 6
        tf_peak = peaks.get("bf16_tflops", 2250.0) / 2.0 / 3.0
        roofline["tensor_kernel"] = {
            "kernel": "linear_tc_kernel (dd_linear_fwd / dd_linear_bwd)", "bound": "hbm", "peak": peak, "unit": "GB/s",
            "tensor_peak_tflops_3xtf32": tf_peak,
            **{k: dict(v, frac=v["algorithmic_gbs"] / peak, tensor_frac=v["tflops_fp32_equivalent"] / tf_peak) for k, v in lin.items()}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        child = _child_json(["--impl", "reference", "--config", args.config, "--phase", args.phase, "--steps", "3", "--warmup", "1"], 900)
        cpu_baseline = child.get("cpu_baseline", child)
    if world == 1 and variants is not None and not args.no_eager:
        # the real bar: the reference itself, eager PyTorch, on this B200 -- in a child process
        # because its module names (Trainer, tools, networks ...) are the product's drop-in names
        torch.cuda.empty_cache()
        for name, tf32 in (("eager_b200", 1), ("eager_b200_tf32_off", 0)):
            variants[name] = _child_json(["--impl", "eager", "--config", args.config, "--phase", args.phase, "--steps", "10", "--warmup", "5",
                                          "--tf32", str(tf32)], 1200)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.phase), "name": args.config, "global_batch": world * BATCH,
                       "parallelism": f"dp{world}",
                       "l2": "per-step working set (colour frames plus GBs of activations) exceeds the 126 MB L2; no flush needed",
                       "encoders": "cuDNN convolutions (TF32 = torch default; channels_last ResNets and Lite-Mono stem); Lite-Mono linear layers: hand-written tcgen05 3xTF32 kernel (fp32 accuracy, csrc/linear_tc.cu); BatchNorm(+GELU), LayerNorm, depth-wise convolutions, XCA attention core, max-pool and layout glue of the encoders, all decoders and the loss path: hand-written kernels at fp32 accuracy",
                       "d_ground": "reference RANSAC prior kept on (host-driven torch ops, SURVEY 8f-1)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clock_info, "collective": collective, "variants": variants}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="litemono_bs32", choices=list(CONFIGS))
    ap.add_argument("--phase", default="fine_tune", choices=["disp_init", "motion_init", "mask_init", "fine_tune"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--tf32", type=int, default=1, help="--impl eager: 1 = torch default (cuDNN TF32), 0 = fp32")
    ap.add_argument("--sample-batch", type=int, default=0, help="--impl reference: triplets per CPU step (default 1-2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip variants.eager_b200 (the reference in eager PyTorch on this GPU)")
    ap.add_argument("--no-variants", action="store_true", help="skip every variant measurement")
    ap.add_argument("--trace-step", default=None, help="write a torch.profiler chrome trace of two steps to this path, no JSON line")
    ap.add_argument("--profile-step", action="store_true", help="cudaProfilerStart/Stop around one step (for ncu), no JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager":
        run_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
