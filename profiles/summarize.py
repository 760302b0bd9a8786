#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: python profiles/summarize.py gpurun_out/X.ncu-rep > profiles/X.txt
(run where ncu is installed; the .ncu-rep itself is scratch and not committed)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64-pipe instructions"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"), ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "L1 local-load sectors (spills)"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "L1 local-store sectors (spills)"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "LSU write-back busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("=" * 100)
        print(d["Kernel Name"])
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"  {label:38s} {d[k]:>18s} {u[k]}")
        st = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    st.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  warp stall reasons (warps stalled per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
