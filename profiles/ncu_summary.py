#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / bench.py cite.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [total_thread_plies]"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in KEYS or h == "Kernel Name":
                d[h] = vals[i] + (" " + units[i] if units[i] else "")
        res.append(d)
    for d in res:
        for k in ["Kernel Name"] + KEYS:
            if k in d:
                print("%-90s %s" % (k, d[k]))
        if len(sys.argv) > 2:
            plies = float(sys.argv[2])
            wi = float(d["smsp__inst_executed.sum"].split()[0])
            lanes = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"].split()[0])
            print("warp_inst_per_thread_ply %.2f   warp_inst_per_warp_ply_if_full %.1f   thread_inst_per_ply %.1f" % (wi / plies, 32 * wi / plies, wi * lanes / plies))
        print()


if __name__ == "__main__":
    main()
