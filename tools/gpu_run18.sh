#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/r2_tc_check4.log 2>&1; grep -i "FAIL\|ALL" gpurun_out/r2_tc_check4.log | head
timeout 300 python tools/tc_bench.py 2>&1 | head -9
timeout 300 python tools/tc_bench.py 1286144 2>&1 | sed -n 2,3p\;6,7p
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_s3dis_g.json 2> gpurun_out/r2_bench_s3dis_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_g.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
for k,v in list(d['kernel_shares'].items())[:6]: print('  ',k, {a:round(b,4) for a,b in v.items()})
PY
