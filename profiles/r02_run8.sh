#!/bin/bash
# round 2, GPU session 8: Pan determinized tree player vs 2 x lowcard (the reference's published experiment, results/mcts_player.log:2-19:
# 64 games, round_limit 100, pts_ratio = pts / 6400 = 0.466 at eer 2.5) over search budgets / deal counts / exploration constants
mkdir -p gpurun_out
L=game-ai_b200/host/GameLauncher
: > gpurun_out/r02_pan_player_sweep.jsonl
run() {
  timeout 900 $L --xml run_config_pan.xml --p1 "gpupan $1" --p2 lowcard --p3 lowcard --ng 64 --threads 8 --seed 5 -q 2> /dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['players'][0]; print(json.dumps({'config':'$1','games':d['games'],'aborted':d['aborted'],'rounds':d['rounds'],'seconds':d['seconds'],'pts':p['pts'],'pts_ratio':p['pts']/(100.0*d['games']),'win':p['win'],'lose':p['lose'],'playouts':float(p['stats']['playouts']),'fallbacks':float(p['stats']['determinization_fallbacks'])}))" | tee -a gpurun_out/r02_pan_player_sweep.jsonl
}
run "move_sim_limit=4096"
run "move_sim_limit=16384"
run "move_sim_limit=65536"
run "move_sim_limit=16384 deals=256"
run "move_sim_limit=16384 explore_exploit_ratio=1.4"
run "move_sim_limit=16384 explore_exploit_ratio=0.35"
run "move_sim_limit=16384 playouts_per_leaf=4"
run "move_sim_limit=16384 playout_depth=100"
run "move_sim_limit=16384 eval_function=num_cards"
run "mode=flat deals=256"
