#!/usr/bin/env python
"""Key counters of one kernel from `ncu -i report.ncu-rep --page raw --csv` (first data row).

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("kernel: %s" % d.get("Kernel Name", "?"))
        for k in KEYS:
            if d.get(k) not in (None, ""):
                print("  %-72s %18s %s" % (k, d[k], u.get(k, "")))
        print("  warp stall reasons (warps per issue-active cycle):")
        st = [(k.split("stalled_")[1].split("_per")[0], float(d[k])) for k in hdr
              if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d.get(k)]
        for name, v in sorted(st, key=lambda kv: -kv[1])[:8]:
            print("    %-24s %.3f" % (name, v))


if __name__ == "__main__":
    main()
