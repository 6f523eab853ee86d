"""Benchmark of the per-step hot path (BASELINE.json): frame-triplets/s of one training step at
192x640, batch 32 per GPU, Lite-Mono depth network, Waymo-shape synthetic triplets, phase fine_tune
(all four networks, every loss term) -- plus the HBM roofline of the fused warp+SSIM kernel and the
reference's CPU path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--phase fine_tune|disp_init] [--impl reference]

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
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "frame_triplets_per_sec_training_step"
UNIT = "triplets/s"
H, W, BATCH = 192, 640, 32
DATASET_KIND = "waymo"
DEPTH_MODEL = "litemono"


def make_opt(batch, local_rank=0):
    import options

    opt = options.DynamoOptions().parse(args=["-d", DATASET_KIND, "--depth_model", DEPTH_MODEL, "--weights_init", "scratch",
                                              "--height", str(H), "--width", str(W), "-b", str(batch)])
    opt.ddp = int(os.environ.get("WORLD_SIZE", 1)) > 1
    opt.local_rank = 0
    opt.cuda_ids = [local_rank]
    return opt


def workload_name(phase):
    return f"litemono Waymo-shape {H}x{W} bs{BATCH} synthetic, phase {phase}"


def algorithmic_bytes(batch, scales, flow_mask):
    """per image and level fwd = 36*P + 4*P/4^s (+16*P/4^s with flow+mask);
    bwd = fwd + 4*P/4^s (+32*P/4^s instead of +16 with flow+mask)."""
    P = H * W
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


def cpu_reference_steps(phase, sample_batch, steps, warmup, seed=1234):
    """The reference's CPU path restated by oracle/ (the reference is a Python tree that cannot travel to
    the GPU box): process_batch + backward + Adam on the host cores, `sample_batch` triplets per step."""
    import networks
    from dd_b200 import synthetic
    from oracle import networks as on

    torch.set_num_threads(os.cpu_count())
    opt = make_opt(sample_batch)
    torch.manual_seed(seed)
    prod = networks.Model(opt)     # host-side construction only: supplies the PyTorch encoders and initial weights
    states = {n: {k: v for k, v in getattr(prod, n).state_dict().items() if not k.startswith("net.")}
              for n in ("depth_dec", "pose_dec", "motion_dec", "motion_mask")}
    om = on.OracleModel(DEPTH_MODEL, opt.scales, opt.frame_ids, prod.depth_enc, prod.pose_enc, prod.motion_enc, states)
    om.train()
    tr = on.OracleTrainer(om, H, W, learning_rate=opt.learning_rate)
    tr.setup_phase(phase)
    tr.step, tr.steps_per_epoch = 100, 100
    batch = synthetic.make_batch(opt, seed)
    from oracle.synth import add_color_pyramid
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        inputs = dict(batch)
        add_color_pyramid(inputs, opt.scales, H, W)      # Trainer.py:729-734 runs on the CPU in the reference
        tr.train_step(inputs)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample_b = 2 if (args.steps + args.warmup) <= 12 else 1
    times, cores = cpu_reference_steps(args.phase, sample_b, args.steps, args.warmup)
    total = sum(times)
    value = sample_b * len(times) / total
    sample = (f"{len(times)} timed steps of {sample_b} triplets each (same synthetic workload and loss terms incl. the RANSAC ground prior; "
              f"bounded sample of the bs{BATCH} step)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(args.phase), "sample_batch": sample_b},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from Trainer import Trainer
    from dd_b200 import _lib as L
    from dd_b200 import functional as Fn
    from dd_b200 import synthetic

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

    opt = make_opt(BATCH, local_rank)
    torch.manual_seed(1234 + rank)
    tr = Trainer(opt)
    tr.setup_phase(args.phase)
    tr.bool_automask = args.phase == "disp_init"
    tr.num_steps_per_epoch = 100
    tr.step = 100                      # loss-weight ramp = 1 (options.py:106-114)
    tr.set_train()

    dev = tr.device
    resident = synthetic.SyntheticTriplets(opt, steps=1, device=dev, seed=1234 + rank, distinct=2).batches
    pinned = synthetic.SyntheticTriplets(opt, steps=1, device=None, seed=1234 + rank, pinned=True, distinct=2).batches

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

    # tensor-core linear kernel (csrc/linear_tc.cu): two extra, untimed-for-the-headline steps with per-call events
    Fn.KERNEL_TIMERS = {"__detail__": True}
    timed(resident, 2, False)
    kt_lin = Fn.KERNEL_TIMERS
    Fn.KERNEL_TIMERS = None
    lin = {}
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

    # variants (not the headline): the Lite-Mono linear layers through torch's fp32 SIMT matmul (the previous default)
    # and through cuBLAS single-pass TF32 (--encoder_tf32_linear, reduced precision) instead of the tcgen05 3xTF32 kernel
    variants = None
    if not args.no_variants:
        from networks.depth_encoder import EncoderLinear
        v_steps = max(3, args.steps // 2)
        variants = {}
        for name, (tf32, mode), note in (
                ("encoder_linear_torch_fp32", (False, "torch"), "options --encoder_linear torch: the Lite-Mono encoder's nn.Linear "
                 "contractions through torch's fp32 matmul (cuBLAS SIMT) instead of csrc/linear_tc.cu"),
                ("encoder_linear_tf32", (True, "tc3x"), "options --encoder_tf32_linear: the same contractions in single-pass TF32 "
                 "(cuBLAS, reduced precision; not parity-grade)")):
            EncoderLinear.tf32, EncoderLinear.mode = tf32, mode
            timed(resident, 2, False)
            ms_v, _, _ = timed(resident, v_steps, False)
            variants[name] = {"value": world * BATCH * v_steps / (ms_v / 1000), "unit": UNIT, "ms_per_step": ms_v / v_steps,
                              "steps": v_steps, "note": note}
        EncoderLinear.tf32, EncoderLinear.mode = False, "tc3x"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * BATCH * args.steps / (ms / 1000)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1000)
    flow_mask = args.phase in ("mask_init", "fine_tune")
    fwd_bytes, bwd_bytes = algorithmic_bytes(BATCH, opt.scales, flow_mask)
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
    roofline = {"bound": "hbm", "kernel": "warp_photo_fwd_kernel (dd_warp_photo_fwd: view synthesis + SSIM/L1 + min, all levels)",
                "achieved": (fwd_bytes / (k_fwd * 1e-3) / 1e9) if k_fwd else None, "peak": peak, "unit": "GB/s",
                "frac": (fwd_bytes / (k_fwd * 1e-3) / 1e9 / peak) if k_fwd else None,
                "traffic": traffic.get(f"warp_photo_fwd_{args.phase}"), "algorithmic_bytes": fwd_bytes, "kernel_ms": k_fwd,
                "peak_source": peak_src,
                "backward": {"kernel": "warp_photo_bwd_kernel", "algorithmic_bytes": bwd_bytes, "kernel_ms": k_bwd,
                             "achieved": (bwd_bytes / (k_bwd * 1e-3) / 1e9) if k_bwd else None,
                             "frac": (bwd_bytes / (k_bwd * 1e-3) / 1e9 / peak) if k_bwd else None,
                             "traffic": traffic.get(f"warp_photo_bwd_{args.phase}")}}
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
        times, cores = cpu_reference_steps(args.phase, 2, 3, 1)
        cpu_baseline = {"value": 2 * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(times)} steps of 2 triplets (bounded sample of the bs{BATCH} step) after 1 warm-up, oracle "
                                  "port of the reference step (all loss terms incl. the RANSAC ground prior)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.phase), "global_batch": world * BATCH, "parallelism": f"dp{world}",
                       "l2": "per-step working set (>= 141 MB of colour frames plus GBs of activations) exceeds the 126 MB L2; no flush needed",
                       "encoders": "PyTorch/cuDNN convolutions (TF32 = torch default, channels_last ResNets); Lite-Mono linear layers: hand-written tcgen05 3xTF32 kernel (fp32 accuracy, csrc/linear_tc.cu); decoders + loss path: hand-written fp32 kernels",
                       "d_ground": "reference RANSAC prior kept on (host-driven torch ops, SURVEY 8f-1)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clock_info, "variants": variants}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--phase", default="fine_tune", choices=["disp_init", "motion_init", "mask_init", "fine_tune"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the opt-in variant measurement (encoder linear layers in TF32)")
    ap.add_argument("--profile-step", action="store_true", help="cudaProfilerStart/Stop around one step (for ncu), no JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
