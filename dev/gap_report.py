"""GPU idle time of a training step from a torch.profiler chrome trace (bench.py --trace-step): gaps between consecutive
kernels on the compute stream, attributed to the kernel that ends the gap and to the CPU op that launched it.
python dev/gap_report.py trace.json [min_gap_us]"""
import json
import sys
from collections import defaultdict


def main(path, min_gap=3.0):
    ev = json.load(open(path))["traceEvents"]
    kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    # main compute stream = the one with the most kernels
    by_stream = defaultdict(list)
    for e in kern:
        by_stream[e.get("args", {}).get("stream", e.get("tid"))].append(e)
    stream, ks = max(by_stream.items(), key=lambda kv: len(kv[1]))
    ks.sort(key=lambda e: e["ts"])
    # launching CPU op by correlation id
    launches = {e["args"]["correlation"]: e for e in ev if e.get("cat") == "cuda_runtime" and "correlation" in e.get("args", {})}
    cpu_ops = sorted((e for e in ev if e.get("cat") == "cpu_op"), key=lambda e: e["ts"])
    span = ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]
    busy = sum(e["dur"] for e in ks)
    print(f"stream {stream}: {len(ks)} kernels, span {span/1e3:.2f} ms, busy {busy/1e3:.2f} ms, idle {(span-busy)/1e3:.2f} ms")
    gaps = defaultdict(lambda: [0, 0.0])
    big = []
    for a, b in zip(ks, ks[1:]):
        gap = b["ts"] - (a["ts"] + a["dur"])
        if gap >= min_gap:
            name = b["name"][:70]
            gaps[name][0] += 1
            gaps[name][1] += gap
            big.append((gap, a["name"][:50], b["name"][:50], b["ts"]))
    tot = sum(v[1] for v in gaps.values())
    print(f"gaps >= {min_gap} us: {tot/1e3:.2f} ms in {sum(v[0] for v in gaps.values())} gaps; by the kernel that follows the gap:")
    for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {v[1]/1e3:7.3f} ms {v[0]:5d}x  {k}")
    print("largest single gaps (us, previous kernel -> next kernel):")
    for g, a, b, ts in sorted(big, reverse=True)[:25]:
        print(f"  {g:8.1f}  {a}  ->  {b}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 3.0)
