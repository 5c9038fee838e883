#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "fused_eval or eval_chunking" > gpurun_out/r2_pytest9a.log 2>&1; tail -25 gpurun_out/r2_pytest9a.log
if grep -q "passed" gpurun_out/r2_pytest9a.log && ! grep -q "failed" gpurun_out/r2_pytest9a.log; then
  for w in room_fwd sema3d_eval; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench_${w}_fused.json 2> gpurun_out/r2_bench_${w}_fused.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${w}_fused.json').read().strip().splitlines()[-1])
print("$w", {k:d.get(k) for k in ('ms_per_step','gpu_launches','parity_rel_err')}, d['e2e']['ms_per_step'])
PY
  done
fi
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest9.log | head -30
