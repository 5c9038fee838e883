#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_check.py > gpurun_out/r2_tc_check.log 2>&1; tail -25 gpurun_out/r2_tc_check.log
timeout 600 python tools/tc_bench.py > gpurun_out/r2_tc_bench.log 2>&1; cat gpurun_out/r2_tc_bench.log
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -40 gpurun_out/r2_pytest2.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_s3dis_b.json 2> gpurun_out/r2_bench_s3dis_b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_b.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','gpu_launches','eager','parity_rel_err'): print(k, d.get(k))
for k,v in d['kernel_shares'].items(): print('  ',k, {a:round(b,4) for a,b in v.items()})
PY
tail -3 gpurun_out/r2_bench_s3dis_b.err
