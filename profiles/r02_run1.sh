#!/bin/bash
# round 2, GPU session 1: parity tests, the bench line (with the player and Pan legs), host-player tuning sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_1.log
tail -5 gpurun_out/r02_gpu_tests_1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench rc=$?"
timeout 900 python profiles/player_sweep.py 6 > gpurun_out/r02_player_sweep_1.jsonl 2> gpurun_out/r02_player_sweep_1.err; echo "sweep rc=$?"
cat gpurun_out/r02_player_sweep_1.jsonl | cut -c1-260
