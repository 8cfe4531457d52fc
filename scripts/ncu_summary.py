#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: one block per captured launch with the figures the
roofline arithmetic needs (duration, DRAM bytes, executed warp instructions, issue-slot utilisation, lanes per
instruction, occupancy limiters, hit rates, top stall reasons).
  python scripts/ncu_summary.py gpurun_out/x.ncu-rep [header line ...] > profiles/rNN_x_ncu_summary.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, notes):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for n in notes:
        print("# " + n)
    print("# report: %s (scratch, not committed)" % rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("\n== %s" % r[col["Kernel Name"]][:150])
        for k in KEYS:
            if k in col:
                print("%-72s %s %s" % (k, r[col[k]], units[col[k]]))
        t_ms = float(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
        inst = float(r[col["smsp__inst_executed.sum"]])
        print("%-72s %.4f   (warp instructions / (148 SMs x 4 schedulers x 1.965 GHz x duration))" %
              ("derived: issue-slot fraction at max clock", inst / (148 * 4 * 1.965e9 * t_ms * 1e-3)))
        st = sorted(((float(r[i]), h) for h, i in col.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")
                     and "not_issued" not in h), reverse=True)
        for v, h in st[:6]:
            print("%-72s %.3f" % (h.replace(STALL, "stall ").replace("_per_issue_active.ratio", " (warps per issue)"), v))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
