#!/bin/bash
# round 2, GPU session 10: whole GPU suite with the staged Pan host-buffer path, default bench line (Pan leg: device-resident value + pinned e2e)
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gpu_tests_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_10.log
tail -14 gpurun_out/r02_gpu_tests_10.log
timeout 900 python bench.py > gpurun_out/r02_bench_final2_1gpu.json 2> gpurun_out/r02_bench_final2_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r02_bench_final2_1gpu.json') if l.startswith('{')][-1])
print(b['value'], b['e2e']['value'], b['gpu_launches'])
print(json.dumps(b['pan'])[:1500])
print({k:v for k,v in b['player'].items() if k not in ('ranks','what')})
PY
timeout 300 python bench.py --workload pan --steps 5 --warmup 2 --no-cpu-baseline 2> /dev/null | cut -c1-600
