#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "local_cloud or pointnet_small or monger" > gpurun_out/r2_pytest20.log 2>&1; tail -4 gpurun_out/r2_pytest20.log; grep -n "^E  " gpurun_out/r2_pytest20.log | head
timeout 900 python bench.py --workload sweep_mat --nodes 10000 --steps 10 --warmup 5 > gpurun_out/r2_bench_sweep_mat.json 2> gpurun_out/r2_bench_sweep_mat.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_sweep_mat.json').read().strip().splitlines()[-1])
    cb=d.get('cpu_baseline') or {}
    print("sweep_mat", d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('gpu_launches'), cb.get('ms_per_step'), cb.get('kind'), cb.get('sample','')[-120:])
except Exception as ex:
    print("failed", ex); print(open('gpurun_out/r2_bench_sweep_mat.err').read()[-1200:])
PY
