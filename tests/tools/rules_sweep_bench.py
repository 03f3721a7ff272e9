#!/usr/bin/env python
"""Config 2 of BASELINE.json in numbers: the LOA rules sweep (move generation, successors, terminal/score) over 1 M
synthetic positions through the C ABI with HOST buffers (H2D + kernel + D2H inside the timed call), for the three device
formulations of the rules, with the reference rules on one host core beside it.  The parity of the same sweep is a test
(tests/test_gpu_rules.py::test_full_size_sweep_1m_positions); this script only times it.
MEASUREMENT SCRIPT, not product: oracle/ serves as corpus generator and as the one-core CPU baseline (like bench.py's
cpu_baseline leg).
usage (GPU box): python tests/tools/rules_sweep_bench.py > gpurun_out/rules_sweep.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gameai_b200 as g
from oracle import loa

port = loa.port()
best = loa.ref() or port
n = 1 << 20
S = np.concatenate([port.gen_reachable(n // 2, 0x10A5EED), port.gen_uniform(n // 2, 0x10A5EED)])
out = {"positions": n, "corpus": "500 k reachable + 500 k uniform (SURVEY 8d)", "gpu": [], "cpu": {}}
eng = g.GpuRollout(0)
for impl, name in ((2, "static sweep"), (1, "line LUT per piece"), (0, "bitboard baseline")):
    eng.set_rules_impl(impl)
    eng.movegen(S[:4096]); eng.successor_hash(S[:4096]); eng.terminal(S[:4096])
    row = {"impl": name}
    for what, fn in (("movegen", eng.movegen), ("successor_hash", eng.successor_hash), ("terminal", eng.terminal)):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); fn(S); ts.append(time.perf_counter() - t0)
        row[what + "_positions_per_sec"] = n / min(ts)
    out["gpu"].append(row)
m = 1 << 16
t0 = time.perf_counter(); best.movegen_batch(S[:m]); dt = time.perf_counter() - t0
out["cpu"]["movegen_positions_per_sec_1core"] = m / dt
t0 = time.perf_counter(); best.terminal_batch(S[:m]); dt = time.perf_counter() - t0
out["cpu"]["terminal_positions_per_sec_1core"] = m / dt
out["cpu"]["kind"] = "reference" if best is not port else "port"
print(json.dumps(out, indent=1))
