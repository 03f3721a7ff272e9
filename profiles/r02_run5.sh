#!/bin/bash
# round 2, GPU session 5: kernel v19 (per-lane copies of the two per-square tables apply reads) parity + bench + ncu; ncu capture of
# the Pan rollout kernel on the bench's own Pan workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_rules.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_gpu_tests_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_5.log
tail -4 gpurun_out/r02_gpu_tests_5.log
timeout 300 python bench.py --steps 6 --warmup 3 --e2e-steps 2 --no-player --no-pan --no-cpu-baseline > gpurun_out/r02_bench_v19_short.json 2> gpurun_out/r02_bench_v19_short.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r02_bench_v19_short.json
timeout 600 python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/plain_v19.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sweep -c 1 -o gpurun_out/prof_v19 -f python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_v19.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --workload pan --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_pan.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pan_rollout -s 1 -c 1 -o gpurun_out/prof_pan -f python bench.py --workload pan --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pan.log 2>&1; echo "ncu pan rc=$?"
tail -2 gpurun_out/plain_pan.log | cut -c1-300
