#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 -k "ecc or shapes" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest6.log | head -30
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_s3dis_c.json 2> gpurun_out/r2_bench_s3dis_c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s3dis_c.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','gpu_launches','parity_rel_err')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
print(json.dumps(d.get('roofline_ecc'), indent=None))
print(d.get('roofline_ecc_scatter_vv'))
print(d.get('roofline_ecc_error'))
PY
tail -3 gpurun_out/r2_bench_s3dis_c.err
