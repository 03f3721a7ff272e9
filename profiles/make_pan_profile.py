#!/usr/bin/env python
"""One `ncu --set full` capture of the Pan rollout kernel + the plain run of the SAME command (bench.py --workload pan) ->
profiles/pan_kernel.json, the constants the Pan leg's roofline object uses (warp-instructions per thread-ply).
usage: python profiles/make_pan_profile.py gpurun_out/prof_pan.ncu-rep gpurun_out/plain_pan.log rNN-tag"""
import csv, json, subprocess, sys, os

rep, plain, tag = sys.argv[1], sys.argv[2], sys.argv[3]
b = json.loads([l for l in open(plain) if l.startswith("{")][-1])
plies_per_launch = b["mean_plies_per_playout"] * b["config"]["deals_per_gpu"] * b["config"]["playouts_per_deal"]
out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.splitlines()))
hdr, vals = rows[0], rows[2]
g = lambda k: float(vals[hdr.index(k)])
wi, lanes = g("smsp__inst_executed.sum"), g("smsp__thread_inst_executed_per_inst_executed.ratio")
d = {"source": "profiles/%s_pan_ncu_summary.txt <- %s (ncu --set full --clock-control none, one launch of %s; %d deals x %d playouts, %d players, depth %d)" % (
        tag.replace("-", "_"), os.path.basename(rep), vals[hdr.index("Kernel Name")].split("(")[0], b["config"]["deals_per_gpu"], b["config"]["playouts_per_deal"],
        b["config"]["players"], b["config"]["max_plies"]),
     "tag": tag, "deals_per_launch": b["config"]["deals_per_gpu"], "thread_plies_per_launch": plies_per_launch, "warp_inst_executed": wi,
     "warp_inst_per_thread_ply": wi / plies_per_launch, "warp_inst_per_warp_ply_if_full": 32 * wi / plies_per_launch, "avg_active_lanes": lanes,
     "thread_inst_per_ply": wi * lanes / plies_per_launch,
     "issue_active_pct": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"), "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), "achieved_occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "registers_per_thread": g("launch__registers_per_thread"), "kernel_ms_under_ncu": g("gpu__time_duration.sum"),
     "dram_bytes_per_launch": g("dram__bytes_read.sum") * 1e6 + g("dram__bytes_write.sum") * 1e6}
json.dump(d, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pan_kernel.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
