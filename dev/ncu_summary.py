"""ncu report -> the compact per-kernel summary kept under profiles/ (same fields as the round-1 summaries).
    python dev/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r02_x_ncu_summary.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^smsp__inst_executed\.sum$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_(fma|alu|lsu|xu|tex)\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$", r"^l1tex__data_pipe_(lsu|tex|tc)_wavefronts(_mem_shared|_mem_lgds)?\.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
        r"^l1tex__t_sector_hit_rate\.pct$", r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit_registers|occupancy_limit_shared_mem|waves_per_multiprocessor)$",
        r"^sm__maximum_warps_per_active_cycle_pct$", r"^sm__cycles_active\.avg$"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("=" * 100)
        print(f"{path} | {r[iname]}")
        stalls = {}
        for h, u, v in zip(hdr, units, r):
            if any(re.search(k, h) for k in KEEP):
                print(f"  {h:86s} {v} {u}")
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h) or \
                re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m:
                try:
                    stalls[m.group(1)] = float(v)
                except ValueError:
                    pass
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
        print("  stalls (warps per issue-active cycle): " + ", ".join(f"{k} {v:.2f}" for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1])
