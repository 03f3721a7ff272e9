#!/bin/bash
# round 2, GPU session 3: parity tests (kernel v17 = v16 + prmt.b32 + eights-plane vote), thread-count variants, full bench line,
# expand_all vs one-leaf-per-selection match, tier-3 match on the real device
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_3.log
tail -4 gpurun_out/r02_gpu_tests_3.log
for T in 896 768 1024; do
  GAI_SWEEP_THREADS=$T timeout 300 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads', $T, 'value %.4e' % d['value'], d['launch'])"
done
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_v17.json 2> gpurun_out/r02_bench_v17.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02_bench_v17.json
L=game-ai_b200/host/GameLauncher
timeout 900 $L --xml run_config_loa.xml --p1 "mcts move_time_limit=0.1 random_seed=1 philox_seed=2" --p2 "mcts move_time_limit=0.1 expand_all=0 gpu_batch=1024 random_seed=3 philox_seed=4" --ng 12 --seed 21 -q | tail -1 | cut -c1-250 > gpurun_out/r02_match_expand_all_as_white.json
timeout 900 $L --xml run_config_loa.xml --p1 "mcts move_time_limit=0.1 expand_all=0 gpu_batch=1024 random_seed=3 philox_seed=4" --p2 "mcts move_time_limit=0.1 random_seed=1 philox_seed=2" --ng 12 --seed 22 -q | tail -1 | cut -c1-250 > gpurun_out/r02_match_expand_all_as_black.json
cat gpurun_out/r02_match_expand_all_as_white.json gpurun_out/r02_match_expand_all_as_black.json
timeout 1500 python profiles/tier3_match.py 64 200 16 > gpurun_out/r02_tier3_match_gpu.json 2> gpurun_out/r02_tier3_match_gpu.err; echo "tier3 rc=$?"
tail -12 gpurun_out/r02_tier3_match_gpu.json
