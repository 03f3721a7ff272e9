#!/bin/bash
# Root-parallel GPU MCTS player (config 4): playouts per second per rank at a fixed 1 s/move budget, 1 rank and N ranks
# (one process per GPU, NCCL root merge per move).  usage (GPU box with N GPUs): bash profiles/player_scaling.sh N
N=${1:-2}
L=game-ai_b200/host/GameLauncher
P1="mcts devices= move_time_limit=1 random_seed=5 philox_seed=6"
run() { # world
  local W=$1; local idf=$(mktemp -u /tmp/ncclid.XXXXXX)
  for r in $(seq 0 $((W-1))); do
    WORLD_SIZE=$W RANK=$r LOCAL_RANK=$r GAI_NCCL_ID_FILE=$idf $L --xml run_config_loa.xml --p1 "$P1" --p2 "random random_seed=7" --ng 1 --rl 24 --seed 11 > /tmp/player_${W}_${r}.log 2>&1 &
  done
  wait
  for r in $(seq 0 $((W-1))); do
    tail -1 /tmp/player_${W}_${r}.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); st = d['players'][0]['stats']
print(json.dumps({'world': $W, 'rank': $r, 'rounds': d['rounds'], 'playouts_per_sec': float(st['playouts_per_sec']), 'plies_per_playout': float(st['plies_per_playout']), 'gpu_busy_fraction': float(st['gpu_busy_fraction']), 'first_move': d['first_move']}))"
  done
}
run 1
if [ "$N" -gt 1 ]; then run $N; fi
