"""Print the metrics that matter from an .ncu-rep (raw page): python dev/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct"]
STALL = "smsp__average_warps_issue_stalled_"


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("=" * 100)
            print(path, "|", r[hdr.index("Kernel Name")][:110])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"  {w:86s} {r[i]} {units[i]}")
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            print("  stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1:])
