#!/bin/bash
# A/B of programmatic dependent launch policies (SPG_PDL = 0 off, 1 all, 2 all but the tensor-core kernels, 3 GEMM+merge only)
mkdir -p gpurun_out
for rep in 1 2; do
for pdl in 0 1 2 3; do
for w in s3dis_train vkitti_train vkitti_eval room_fwd sema3d_eval; do
SPG_PDL=$pdl timeout 600 python bench.py --workload $w --steps 40 --warmup 8 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/pdl${pdl}_$w.json 2> gpurun_out/pdl${pdl}_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/pdl${pdl}_$w.json').read().strip().splitlines()[-1])
    print('pdl=$pdl', '$w', round(d['ms_per_step'],4), 'e2e', round(d['e2e'].get('ms_per_step',0),4))
except Exception as ex:
    print('pdl=$pdl $w FAILED', ex); print(open('gpurun_out/pdl${pdl}_$w.err').read()[-1500:])
PY
done; done; done
