#!/usr/bin/env python
"""Latency of small leaf batches (the MCTS player's case) through gai_loa_rollout: kernel and end-to-end milliseconds.
usage (GPU box): python profiles/small_batch.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gameai_b200 as g

eng = g.GpuRollout(0)
S = eng.gen_reachable(16384, 0x5EED)            # the device corpus generator (k_gen_reachable); nothing from oracle/ here
out = []
for n, ppl in ((256, 32), (1024, 32), (1024, 128), (4096, 32), (16384, 32)):
    eng.rollout(S[:n], ppl, 50000, 1, 0)
    best = None
    for rep in range(5):
        c, st = eng.rollout(S[:n], ppl, 50000, 1, rep * n * ppl)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    out.append({"leaves": n, "playouts_per_leaf": ppl, "kernel_ms": round(best["kernel_ms"], 3), "e2e_ms": round(best["total_ms"], 3),
                "playouts_per_sec": best["playouts"] / (best["total_ms"] * 1e-3)})
print(json.dumps({"spread": "GAI_NO_SPREAD" not in os.environ, "batches": out}, indent=1))
