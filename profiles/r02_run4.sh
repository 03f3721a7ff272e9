#!/bin/bash
# round 2, GPU session 4: the whole GPU suite (session 3 stopped at its first test: the patched reference player's budget was
# far too large for its own host code), kernel v18 (LUT address = PRMT + one IDP.4A) bench + ncu capture
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02_gpu_tests_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_4.log
tail -25 gpurun_out/r02_gpu_tests_4.log
timeout 300 python bench.py --steps 6 --warmup 3 --e2e-steps 2 --no-player --no-pan --no-cpu-baseline > gpurun_out/r02_bench_v18_short.json 2> gpurun_out/r02_bench_v18_short.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r02_bench_v18_short.json
timeout 600 python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/plain_v18.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sweep -c 1 -o gpurun_out/prof_v18 -f python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_v18.log 2>&1; echo "ncu rc=$?"
