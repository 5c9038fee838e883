#!/bin/bash
# A/B of the training-step launch policy: programmatic dependent launch in the forward phase on / off
mkdir -p gpurun_out
for rep in 1 2 3; do
for f in 1 0; do
for w in s3dis_train vkitti_train; do
SPG_PDL_TRAIN_FWD=$f timeout 600 python bench.py --workload $w --steps 40 --warmup 8 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/pdlf${f}_$w.json 2> gpurun_out/pdlf${f}_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/pdlf${f}_$w.json').read().strip().splitlines()[-1])
    print('fwd_pdl=$f', '$w', round(d['ms_per_step'],4), 'e2e', round(d['e2e'].get('ms_per_step',0),4))
except Exception as ex:
    print('fwd_pdl=$f $w FAILED', ex); print(open('gpurun_out/pdlf${f}_$w.err').read()[-1500:])
PY
done; done; done
