"""MEASUREMENT SCRIPT, not product: one Pan rollout launch (3 players) for ncu; oracle/ only generates the corpus of deals.
usage (GPU box): python profiles/pan_one.py [max_plies]"""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, gameai_b200 as g
from oracle import pan
eng = g.GpuRollout(0)
npl, cap = 3, int(sys.argv[1]) if len(sys.argv) > 1 else 50
S = pan.port().gen_reachable(npl, 1 << 18, 0x10A5EED + npl, 60)
eng.pan_rollout(npl, S[:4096], 16, cap, 1, 1, 0)
r, st = eng.pan_rollout(npl, S, 16, cap, 1, 1, 0)
print(cap, st)
