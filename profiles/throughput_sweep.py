#!/usr/bin/env python
"""MEASUREMENT SCRIPT: the batch sizes of SURVEY 8d through bench.py (device-resident value + host-buffer e2e), one JSON line each.
usage (GPU box): python profiles/throughput_sweep.py > gpurun_out/throughput_sweep.jsonl"""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for leaves in (65536, 262144, 1048576):
    for ppl in (1, 16):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--leaves", str(leaves), "--ppl", str(ppl), "--steps", "5", "--warmup", "3", "--e2e-steps", "3",
                            "--no-player", "--no-pan", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        print(json.dumps({"leaves": leaves, "playouts_per_leaf": ppl, "playouts_per_sec": d["value"], "e2e_playouts_per_sec": d["e2e"]["value"],
                          "ms_per_launch": d["ms_per_step"], "mean_plies": d["mean_plies_per_playout"], "clocks": d["clocks"]}), flush=True)
