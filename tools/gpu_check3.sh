#!/bin/bash
# device graph build + pipelined inference upload: parity subset, then the e2e numbers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_bf16.py tests/test_main_unchanged.py -m gpu -q -x --timeout 900 > gpurun_out/chk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/chk3_pytest.log
tail -15 gpurun_out/chk3_pytest.log
for spec in "s3dis_train 1024 40" "sema3d_eval 20000 20" "vkitti_eval 8192 30" "room_fwd 1536 30"; do set -- $spec
timeout 600 python bench.py --workload $1 --nodes $2 --steps $3 --warmup 8 --no-roofline --no-cpu-baseline > gpurun_out/chk3_$1.json 2> gpurun_out/chk3_$1.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/chk3_$1.json').read().strip().splitlines()[-1])
    print('$1', round(d['ms_per_step'],4), 'e2e', round(d['e2e'].get('ms_per_step',0),4), 'h2d', d['e2e']['h2d_bytes_per_step'], 'parity', d.get('parity_rel_err'))
except Exception as ex:
    print('$1 FAILED', ex); print(open('gpurun_out/chk3_$1.err').read()[-2500:])
PY
done
