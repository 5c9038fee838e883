#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 2>&1 | grep -i "FAIL\|ALL" 
timeout 300 python tools/tc_bench.py > gpurun_out/r2_tc_bench3.log 2>&1; cat gpurun_out/r2_tc_bench3.log
python tools/fused_one.py 18248; python tools/fused_one.py 942
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench_s3dis_e.json 2> gpurun_out/r2_bench_s3dis_e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_e.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','gpu_launches','parity_rel_err')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
PY
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest15.log | head -30
