#!/usr/bin/env python
"""ncu_pick.py report.ncu-rep [substring ...] -- the named metrics of every kernel in an ncu report (raw page), one per line."""
import csv
import io
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads",
           "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads", "sm__inst_executed_pipe_fp64.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "lts__t_sector_hit_rate.pct",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[kcol][:90])
        for h, u, v in zip(hdr, units, r):
            if (not pats and h in DEFAULT) or any(p in h for p in pats):
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main()
