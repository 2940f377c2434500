#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture (.ncu-rep) into the text kept under profiles/.

usage: tools/ncu_summary.py REP "header line" > profiles/rNN_ncu_<kernel>_summary.txt
Runs `ncu -i REP --page raw --csv` here (no GPU needed) and keeps the metrics the design discussion uses."""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
        "sm__sass_thread_inst_executed_op_integer_pred_on.sum", "smsp__inst_executed_op_branch.sum")


def main():
    rep, header = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    print(header)
    kn = names.index("Kernel Name") if "Kernel Name" in names else None
    if kn is not None:
        print("kernel", vals[kn])
    for n, u, v in sorted(zip(names, units, vals)):
        if any(n.startswith(k) for k in KEEP):
            print(n, u, v)


if __name__ == "__main__":
    main()
