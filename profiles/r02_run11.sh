#!/bin/bash
# round 2, GPU session 11 (two GPUs): the driver's N=2 command with the final code (two trees per rank in the player leg, staged Pan path)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_final_2gpu.json 2> gpurun_out/r02_bench_final_2gpu.err; echo "bench N=2 rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r02_bench_final_2gpu.json') if l.startswith('{')][-1])
print(b['value'], b['e2e']['value'], b['n_gpus'])
p=b['player']; print({k:v for k,v in p.items() if k not in ('ranks','what')})
for r in p['ranks']: print({k:v for k,v in r.items() if k not in ('moves','playouts_per_move_histogram')})
print(b['pan']['value'], b['pan']['e2e']['value'])
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
