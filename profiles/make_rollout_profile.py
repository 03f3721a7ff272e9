#!/usr/bin/env python
"""Turn one `ncu --set full` capture of the rollout kernel + the plain (un-profiled) run of the SAME command into
profiles/rollout_kernel.json, the constants bench.py's roofline object uses (warp-instructions per thread-ply).
usage: python profiles/make_rollout_profile.py gpurun_out/prof.ncu-rep gpurun_out/plain.log rNN-tag"""
import csv, json, subprocess, sys, os

rep, plain, tag = sys.argv[1], sys.argv[2], sys.argv[3]
line = [l for l in open(plain) if l.startswith("{")][-1]
b = json.loads(line)
plies_per_launch = b["mean_plies_per_playout"] * b["playouts_per_step"]
out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.splitlines()))
hdr, vals = rows[0], rows[2]
g = lambda k: float(vals[hdr.index(k)])
wi = g("smsp__inst_executed.sum")
lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
d = {
    "source": "profiles/%s_ncu_summary.txt <- %s (ncu --set full --clock-control none, one launch of %s; %d leaves x %d playouts)" % (
        tag.replace("-", "_"), os.path.basename(rep), vals[hdr.index("Kernel Name")].split("(")[0], b["config"]["leaves_per_gpu"], b["config"]["playouts_per_leaf"]),
    "tag": tag,
    "leaves_per_launch": b["config"]["leaves_per_gpu"],
    "thread_plies_per_launch": plies_per_launch,
    "warp_inst_executed": wi,
    "warp_inst_per_thread_ply": wi / plies_per_launch,
    "warp_inst_per_warp_ply_if_full": 32 * wi / plies_per_launch,
    "avg_active_lanes": lanes,
    "thread_inst_per_ply": wi * lanes / plies_per_launch,
    "issue_active_pct": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "achieved_occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": g("launch__registers_per_thread"),
    "kernel_ms_under_ncu": g("gpu__time_duration.sum"),
    "dram_bytes_per_launch": g("dram__bytes_read.sum") * 1e6 + g("dram__bytes_write.sum"),
    "shared_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflict_wavefronts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
}
json.dump(d, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rollout_kernel.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
