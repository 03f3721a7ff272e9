#!/bin/bash
# round 2, GPU session 9: final numbers for kernel v19 -- the default bench line, its launch list, the batch-size sweep, the small-batch
# latencies; the longer tier-3 match (GPU player vs the reference's MC::Player, control, both vs random); two more Pan configurations
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/r02_bench_final_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v19.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_ll_v19.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 python profiles/throughput_sweep.py > gpurun_out/r02_throughput_sweep.jsonl 2> gpurun_out/r02_throughput_sweep.err; echo "sweep rc=$?"; cut -c1-200 gpurun_out/r02_throughput_sweep.jsonl
timeout 300 python profiles/small_batch.py > gpurun_out/r02_small_batch_latency.json 2>&1; echo "small batch rc=$?"; tail -3 gpurun_out/r02_small_batch_latency.json | cut -c1-300
L=game-ai_b200/host/GameLauncher
run() {
  timeout 900 $L --xml run_config_pan.xml --p1 "gpupan $1" --p2 lowcard --p3 lowcard --ng 64 --threads 8 --seed 5 -q 2> /dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['players'][0]; print(json.dumps({'config':'$1','games':d['games'],'aborted':d['aborted'],'rounds':d['rounds'],'seconds':d['seconds'],'pts':p['pts'],'pts_ratio':p['pts']/(100.0*d['games']),'win':p['win'],'lose':p['lose'],'playouts':float(p['stats']['playouts']),'fallbacks':float(p['stats']['determinization_fallbacks'])}))" | tee -a gpurun_out/r02_pan_player_sweep2.jsonl
}
: > gpurun_out/r02_pan_player_sweep2.jsonl
run "move_sim_limit=16384 eval_function=num_cards playout_depth=100"
run "move_sim_limit=16384 playout_depth=1000"
run "move_sim_limit=16384 eval_function=num_cards explore_exploit_ratio=0.35"
timeout 2400 python tests/tools/tier3_match.py 64 200 16 > gpurun_out/r02_tier3_match.json 2> gpurun_out/r02_tier3_match.err; echo "tier3 rc=$?"
tail -14 gpurun_out/r02_tier3_match.json
