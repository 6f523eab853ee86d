"""Hot spots of one kernel from an .ncu-rep source page (SASS view): samples by region.
python dev/ncu_hot.py file.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, regex, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = [r for r in rows[2:] if len(r) > max(i_s, i_ex) and r[i_s].isdigit()]
    tot_s = sum(int(r[i_s]) for r in body)
    tot_e = sum(int(r[i_ex]) for r in body)
    print(f"{rows[0][1]}: {len(body)} SASS instructions, {tot_s} samples, {tot_e} warp-instructions executed")
    # windowed view: consecutive chunks of 32 instructions
    print("--- by 48-instruction window: idx  samples%  exec%  first opcode(s)")
    W = 48
    for k in range(0, len(body), W):
        ch = body[k:k + W]
        s = sum(int(r[i_s]) for r in ch)
        e = sum(int(r[i_ex]) for r in ch)
        if s * 100 / max(tot_s, 1) >= 1.0:
            ops = {}
            for r in ch:
                op = r[i_src].split()[0] if not r[i_src].strip().startswith("@") else r[i_src].split()[1]
                op = op.split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            desc = " ".join(f"{o}:{n}" for o, n in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
            print(f"{k:6d} {100*s/tot_s:6.1f}% {100*e/tot_e:6.1f}%  {desc}")
    print(f"--- top {top} instructions by samples")
    order = sorted(range(len(body)), key=lambda j: -int(body[j][i_s]))[:top]
    for j in sorted(order):
        r = body[j]
        print(f"{j:6d} {100*int(r[i_s])/tot_s:5.2f}% ex {int(r[i_ex]):9d}  {r[i_src].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
