#!/bin/bash
# round 2, GPU session 2: parity tests with the chi-gated kernel (v16), bench, ncu capture + launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_2.log
tail -4 gpurun_out/r02_gpu_tests_2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-player --no-pan --no-cpu-baseline > gpurun_out/r02_bench_v16.json 2> gpurun_out/r02_bench_v16.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r02_bench_v16.json
# plain run of the profiled command, then ncu on the same command
timeout 600 python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/plain_v16.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sweep -c 1 -o gpurun_out/prof_v16 -f python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_v16.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v16.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_ll_v16.log 2>&1; echo "ncu launch list rc=$?"
