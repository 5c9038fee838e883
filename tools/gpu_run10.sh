#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "ragged" 2>&1 | tail -3
for w in room_fwd sema3d_eval; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2_bench_${w}_ks.json 2> gpurun_out/r2_bench_${w}_ks.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${w}_ks.json').read().strip().splitlines()[-1])
print("$w", d['ms_per_step'])
for k,v in d.get('kernel_shares',{}).items(): print('   ',k, {a:round(b,4) for a,b in v.items()})
PY
done
