#!/bin/bash
# round 2, GPU session 7: the config-4 player with 1 / 2 / 3 root-parallel trees on ONE GPU (one host thread, context and stream per
# tree), 10 move decisions at 1 s each from the opening vs a fixed-seed random mover; Pan: the determinized tree player vs 2 x lowcard
mkdir -p gpurun_out
L=game-ai_b200/host/GameLauncher
: > gpurun_out/r02_player_trees_per_gpu.jsonl
for D in "0" "0,0" "0,0,0" "0,0,0,0"; do
  for B in 512 1024; do
    timeout 120 $L --xml run_config_loa.xml --p1 "random random_seed=7" --p2 "mcts devices=$D gpu_batch=$B move_time_limit=1 random_seed=5 philox_seed=6" --ng 1 --rl 20 --seed 11 -q 2> /dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['players'][1]['stats']; print(json.dumps({'devices':'$D','gpu_batch':$B,'playouts_per_sec':float(s['playouts_per_sec']),'gpu_busy_fraction':float(s['gpu_busy_fraction']),'selections_per_sec':float(s['selections_per_sec']),'gpu_batches':float(s['gpu_batches']),'runs_per_move':s['runs_per_move']}))" | tee -a gpurun_out/r02_player_trees_per_gpu.jsonl | cut -c1-220
  done
done
timeout 1200 $L --xml run_config_pan.xml --p1 gpupan --p2 lowcard --p3 lowcard --ng 64 --threads 8 --seed 5 -q 2> gpurun_out/r02_pan_match_tree.err | tail -1 > gpurun_out/r02_pan_match_tree.json; echo "pan tree match rc=$?"
cut -c1-700 gpurun_out/r02_pan_match_tree.json
timeout 1200 $L --xml run_config_pan.xml --p1 "gpupan mode=flat" --p2 lowcard --p3 lowcard --ng 64 --threads 8 --seed 5 -q 2> gpurun_out/r02_pan_match_flat.err | tail -1 > gpurun_out/r02_pan_match_flat.json; echo "pan flat match rc=$?"
cut -c1-700 gpurun_out/r02_pan_match_flat.json
