#!/bin/bash
# round 2, GPU session 3: parity tests (kernel v17 = v16 + prmt.b32 + eights-plane vote), thread-count variants, full bench line,
# ncu capture + launch list of v17
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_3.log
tail -4 gpurun_out/r02_gpu_tests_3.log
for T in 896 768 1024; do
  GAI_SWEEP_THREADS=$T timeout 300 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads', $T, 'value %.4e' % d['value'], d['launch'])"
done
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_v17.json 2> gpurun_out/r02_bench_v17.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02_bench_v17.json
timeout 600 python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/plain_v17.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sweep -c 1 -o gpurun_out/prof_v17 -f python bench.py --leaves 262144 --ppl 16 --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_v17.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v17.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-player --no-pan --no-cpu-baseline > gpurun_out/ncu_ll_v17.log 2>&1; echo "ncu launch list rc=$?"
