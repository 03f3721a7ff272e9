#!/bin/bash
# round 2, GPU session 12 (eight GPUs): the driver's N=8 and N=4 commands with the final code
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_final_${N}gpu.json 2> gpurun_out/r02_bench_final_${N}gpu.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/r02_bench_final_${N}gpu.json') if l.startswith('{')][-1])
print(b['value'], b['e2e']['value'], b['n_gpus'], b['clocks'])
p=b['player']; print({k:v for k,v in p.items() if k not in ('ranks','what')})
for r in p['ranks']: print({k:v for k,v in r.items() if k in ('rank','rc','playouts_per_sec','gpu_busy_fraction','allreduce_us_per_move','error')})
print(b['pan']['value'], b['pan']['e2e']['value'])
PY
done
