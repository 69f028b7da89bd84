"""Summarise ncu artefacts into small text files for profiles/ (run in the build container).

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep   > profiles/r01_xxx.txt
    python tools/ncu_summary.py list  gpurun_out/launches.csv   > profiles/r01_launches_xxx.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("  %-82s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        print()


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 8]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(dict)
    for r in rows[1:]:
        try:
            per[(r[ii], r[ki])][r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    agg = defaultdict(lambda: [0, 0.0])
    for (_, k), m in per.items():
        agg[k][0] += 1
        agg[k][1] += m.get("gpu__time_duration.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("launches: %d   total device time (serialised, cold cache): %.1f us" % (sum(a[0] for a in agg.values()), tot / 1e3))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("  share %.4f  n=%4d  avg %9.2f us  %s" % (a[1] / tot, a[0], a[1] / a[0] / 1e3, k[:150]))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
