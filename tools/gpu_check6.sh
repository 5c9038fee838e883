#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x --timeout 600 -k "sema3d or eval_graph" > gpurun_out/chk6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/chk6_pytest.log
tail -3 gpurun_out/chk6_pytest.log
for spec in "room_fwd 1536 30" "vkitti_eval 8192 30" "sema3d_eval 20000 20"; do set -- $spec
timeout 600 python bench.py --workload $1 --nodes $2 --steps $3 --warmup 8 --no-cpu-baseline --no-parity > gpurun_out/chk6_$1.json 2> gpurun_out/chk6_$1.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/chk6_$1.json').read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    print('$1', round(d['ms_per_step'],4), 'e2e', round(d['e2e'].get('ms_per_step',0),4), 'roofline', r.get('kernel'), r.get('achieved'), r.get('frac'), r.get('share_of_step'), r.get('traffic'))
except Exception as ex:
    print('$1 FAILED', ex); print(open('gpurun_out/chk6_$1.err').read()[-2500:])
PY
done
