"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_step_launches.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    ik, iv, iu = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if len(r) <= iv:
            continue
        val = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
        name = re.sub(r"\(.*$", "", r[ik])
        name = name if len(name) < 90 else name[:87] + "..."
        agg[name][0] += 1
        agg[name][1] += us
        total += us
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {total/1000:.2f} ms summed kernel time (cold-cache, serialised)")
    print(f"{'share':>7} {'ms':>9} {'count':>6}  kernel")
    ours = 0.0
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "dd::" in name:
            ours += us
        print(f"{100*us/total:6.2f}% {us/1000:9.3f} {n:6d}  {name}")
    print(f"# hand-written kernels (dd::*): {ours/1000:.2f} ms = {100*ours/total:.1f}% of summed kernel time")


if __name__ == "__main__":
    main(sys.argv[1])
