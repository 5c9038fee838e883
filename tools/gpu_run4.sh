#!/bin/bash
# 4 GPUs: scene-parallel bench with the fused all-reduce kernel (world > 2 check), bounded
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 30 --warmup 5 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/r2_bench_4gpu_s3dis_train.json 2> gpurun_out/r2_bench_4gpu.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_4gpu_s3dis_train.json').read().strip().splitlines()[-1])
    print("4gpu", d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
except Exception as ex:
    print("bench failed", ex); print(open('gpurun_out/r2_bench_4gpu.err').read()[-1500:])
PY
