#!/usr/bin/env python
"""Host-side tuning of the GPU MCTS player (config 4) on one GPU: playouts/s, selections/s and GPU busy fraction at 1 s per
move from the opening for several batch sizes / pipeline depths / playouts per leaf.  MEASUREMENT SCRIPT.
usage (GPU box): python profiles/player_sweep.py [moves] > gpurun_out/player_sweep.jsonl"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 6
configs = sys.argv[2:] or [
    "expand_all=1 gpu_batch=1024 pipeline_depth=2 playouts_per_leaf=32",
    "expand_all=1 gpu_batch=1024 pipeline_depth=3 playouts_per_leaf=32",
    "expand_all=1 gpu_batch=512 pipeline_depth=2 playouts_per_leaf=32",
    "expand_all=1 gpu_batch=512 pipeline_depth=3 playouts_per_leaf=32",
    "expand_all=1 gpu_batch=256 pipeline_depth=3 playouts_per_leaf=32",
    "expand_all=1 gpu_batch=2048 pipeline_depth=2 playouts_per_leaf=16",
    "expand_all=1 gpu_batch=1024 pipeline_depth=3 playouts_per_leaf=16",
    "expand_all=1 gpu_batch=2048 pipeline_depth=3 playouts_per_leaf=8",
    "expand_all=0 gpu_batch=1024 pipeline_depth=2 playouts_per_leaf=32",
    "expand_all=0 gpu_batch=4096 pipeline_depth=3 playouts_per_leaf=32",
]
for c in configs:
    r = subprocess.run([L, "--xml", os.path.join(ROOT, "run_config_loa.xml"), "--p1", "random random_seed=7", "--p2", "mcts move_time_limit=1 random_seed=5 philox_seed=6 " + c,
                        "--ng", "1", "--rl", str(2 * moves), "--seed", "11"], capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        print(json.dumps({"config": c, "error": r.stderr[-300:]}), flush=True)
        continue
    st = json.loads(r.stdout.strip().splitlines()[-1])["players"][1]["stats"]
    print(json.dumps({"config": c, "playouts_per_sec": float(st["playouts_per_sec"]), "selections_per_sec": float(st["selections_per_sec"]),
                      "gpu_busy_fraction": float(st["gpu_busy_fraction"]), "plies_per_playout": float(st["plies_per_playout"]),
                      "gpu_batches": float(st["gpu_batches"]), "pending_collisions": float(st["pending_collisions"]), "runs_per_move": st["runs_per_move"]}), flush=True)
