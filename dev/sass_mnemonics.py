"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses (cuobjdump -sass of the built .so).
python dev/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200", "dd_b200", "libdynamo_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ELECT", "TLD4", "LDGSTS", "FFMA2", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur:
            m = re.search(r"^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                kernels[cur][m.group(1).split(".")[0]] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("SASS mnemonics per kernel of dynamo-depth_b200/dd_b200/libdynamo_b200.so (cuobjdump -sass; round 2, final build)")
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, TLD4 = tex2Dgather, UTCBAR = tcgen05.commit, SYNCS = mbarrier, "
          "LDGSTS = cp.async, FFMA2 = packed fp32 FMA, UTMALDG = cp.async.bulk.tensor (tensor-map TMA: the token tiles of the XCA kernels); "
          "the tensor-core kernels stage their operands through registers (hi / lo split) or as pre-arranged images by UBLKCP")
    print()
    seen = set()
    for (name, cnt), dem in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*$", "", dem)
        short = re.sub(r"linear_tc_kernel<(\w+), (\w+), \d+, (\w+)>", r"linear_tc_kernel<\1, \2, NB32 = 1..8, \3>", short)
        short = re.sub(r"linear_prep_b_kernel<(\w+), \d+>", r"linear_prep_b_kernel<\1, NB32 = 1..8>", short)
        if short in seen:
            continue
        seen.add(short)
        hits = " ".join(f"{k}={cnt[k]}" for k in WATCH if cnt.get(k))
        print(f"{short[:100]:100s} {sum(cnt.values()):6d} instr  {hits}")


if __name__ == "__main__":
    main()
