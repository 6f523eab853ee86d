"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals (the format of profiles/*_step_launches_*.txt)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, title=""):
    rows = list(csv.reader(l for l in open(path, errors="ignore") if l.startswith('"')))
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= ival:
            continue
        v = float(r[ival].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit].lower().replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1e-6)
        name = re.sub(r"\(.*$", "", r[iname])[:110]
        tot[name] += v
        cnt[name] += 1
    print(title or f"ncu launch list {path}, summed by kernel")
    print(f"total {sum(tot.values()):.2f} ms over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:8.2f} ms {cnt[k]:5d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
