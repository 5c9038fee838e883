#!/bin/bash
mkdir -p gpurun_out
for m in 1 0; do
echo "== SPG_FUSED_A_TMEM=$m"
SPG_FUSED_A_TMEM=$m timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "fused_eval or eval_chunking or ragged" 2>&1 | tail -4
SPG_FUSED_A_TMEM=$m timeout 600 python bench.py --workload sema3d_eval --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2_bench_sema_atm$m.json 2> gpurun_out/r2_bench_sema_atm$m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_sema_atm$m.json').read().strip().splitlines()[-1])
print("sema3d_eval", d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in list(d.get('kernel_shares',{}).items())[:3]})
PY
done
