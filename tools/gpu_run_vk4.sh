#!/bin/bash
mkdir -p gpurun_out
for w in vkitti_eval vkitti_train s3dis_train; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --workload $w --steps 30 --warmup 8 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/r2_bench_4gpu_$w.json 2> gpurun_out/r2_bench_4gpu_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_4gpu_$w.json').read().strip().splitlines()[-1])
    print("4gpu $w", d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['dtype'])
except Exception as ex:
    print("bench failed", ex); print(open('gpurun_out/r2_bench_4gpu_$w.err').read()[-1500:])
PY
done
