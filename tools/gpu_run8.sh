#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/grad_diag.py 2>&1 | grep -v "stn.convs\|stn.fcs" > gpurun_out/r2_grad_diag2.log; grep "ptn\.\|parameter" gpurun_out/r2_grad_diag2.log
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest8.log | head -30
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench_s3dis_d.json 2> gpurun_out/r2_bench_s3dis_d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_d.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','gpu_launches','parity_rel_err')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
PY
