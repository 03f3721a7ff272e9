#!/bin/bash
# round 2, GPU session 6 (two GPUs): the two-rank NCCL player test, the driver's N=2 bench command (synthetic + Pan + the config-4
# player leg with the per-move NCCL merge inside the measured loop) and the reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_gpu_tests_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_multi.log
tail -3 gpurun_out/r02_gpu_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_v19_2gpu.json 2> gpurun_out/r02_bench_v19_2gpu.err; echo "bench N=2 rc=$?"
cut -c1-300 gpurun_out/r02_bench_v19_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_2gpu.json 2> gpurun_out/r02_bench_ref_2gpu.err; echo "reference arm N=2 rc=$?"
cut -c1-300 gpurun_out/r02_bench_ref_2gpu.json
