#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_bf16.py -m gpu -q --timeout 200 > gpurun_out/r2_pytest19.log 2>&1; tail -5 gpurun_out/r2_pytest19.log; grep -n "^E  " gpurun_out/r2_pytest19.log | head -10
timeout 300 python bench.py --workload vkitti_eval --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_vkitti_eval.json 2> gpurun_out/r2_bench_vkitti_eval.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_vkitti_eval.json').read().strip().splitlines()[-1])
    print("vkitti_eval", {k:d.get(k) for k in ('ms_per_step','value','gpu_launches','parity_rel_err','dtype')}, d['e2e']['ms_per_step'])
    for k,v in list(d.get('kernel_shares',{}).items())[:4]: print('  ',k, {a:round(b,4) for a,b in v.items()})
except Exception as ex:
    print("failed", ex); print(open('gpurun_out/r2_bench_vkitti_eval.err').read()[-1200:])
PY
