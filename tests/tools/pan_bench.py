#!/usr/bin/env python
"""Pan determinized-rollout throughput on one B200 (config 5 of BASELINE.json), with the reference rules timed beside it.
MEASUREMENT SCRIPT, not product: oracle/ serves as corpus generator, as the checker of the first 2048 deals and as the
CPU baseline beside the GPU number (like bench.py's cpu_baseline leg).
usage (GPU box): python tests/tools/pan_bench.py > gpurun_out/pan_bench.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gameai_b200 as g
from oracle import pan

eng = g.GpuRollout(0)
po = pan.best()
out = []
for npl, cap in ((2, 50), (3, 50), (4, 50), (3, 100000)):
    n, ppl = 1 << 18, 16
    S = pan.port().gen_reachable(npl, n, 0x10A5EED + npl, 60)           # sampled (determinized) deals, mid-game
    eng.pan_rollout(npl, S[:4096], ppl, cap, 1, 1, 0)                    # warm-up
    best = None
    for rep in range(3):
        r, st = eng.pan_rollout(npl, S, ppl, cap, 1, 1, 0)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    t0 = time.perf_counter(); orr, oplies = po.playout_batch(npl, S[:2048], ppl, cap, 1, 1, 0); dt = time.perf_counter() - t0
    assert (orr == r[:2048]).all()
    out.append({"players": npl, "max_plies": cap, "deals": n, "playouts_per_deal": ppl, "kernel_ms": best["kernel_ms"], "e2e_ms": best["total_ms"],
                "gpu_playouts_per_sec": best["playouts"] / (best["kernel_ms"] * 1e-3), "gpu_e2e_playouts_per_sec": best["playouts"] / (best["total_ms"] * 1e-3),
                "mean_plies": best["plies"] / best["playouts"], "gpu_plies_per_sec": best["plies"] / (best["kernel_ms"] * 1e-3),
                "cpu_1core_playouts_per_sec": 2048 * ppl / dt, "cpu_kind": po.kind, "checked_vs_oracle_leaves": 2048})
print(json.dumps(out, indent=1))
