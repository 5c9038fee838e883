#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 2>&1 | grep -i "fused\|FAIL\|ALL" > gpurun_out/r2_tc_check3.log; cat gpurun_out/r2_tc_check3.log
for f in 1 0; do
SPG_FUSED_BNBWD=$f python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_s3dis_f$f.json 2> gpurun_out/r2_bench_s3dis_f$f.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_f$f.json').read().strip().splitlines()[-1])
print("FUSED_BNBWD=$f", {k:d.get(k) for k in ('ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
for k,v in list(d['kernel_shares'].items())[:9]: print('  ',k, {a:round(b,4) for a,b in v.items()})
PY
done
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest3.log | head -30
