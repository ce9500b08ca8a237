#!/usr/bin/env python3
"""profiles/ncu_summary.py — print the metrics we track from an .ncu-rep (run here, no GPU needed).
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [substring ...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__throughput.avg.pct",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled", "smsp__issue_active.avg.pct", "l1tex__throughput.avg.pct", "lts__t_sector_hit_rate",
        "smsp__inst_executed_op_shared", "sm__pipe_fp64", "smsp__pcsamp_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(k.strip() in h for k in KEYS + extra) and r[i] not in ("", "0"):
                print(f"  {h:95s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
