#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 2>&1 | grep -i "FAIL\|ALL"
timeout 300 python tools/tc_bench.py 2>&1 | head -9
timeout 300 python tools/tc_bench.py 1286144 2>&1 | sed -n 2,3p\;6,7p
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-parity > gpurun_out/r2_bench_s3dis_f.json 2> gpurun_out/r2_bench_s3dis_f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_f.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
PY
timeout 900 python bench.py --workload sweep_vv --nodes 100000 --steps 5 --warmup 2 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench_sweep_vv_100000.json 2> gpurun_out/r2_bench_sweep_vv_100000.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_sweep_vv_100000.json').read().strip().splitlines()[-1])
    print("sweep_vv 100000", {k:d.get(k) for k in ('ms_per_step','value','gpu_launches','cuda_graph')}, d['e2e']['ms_per_step'])
except Exception as ex:
    print("sweep failed", ex); print(open('gpurun_out/r2_bench_sweep_vv_100000.err').read()[-800:])
PY
