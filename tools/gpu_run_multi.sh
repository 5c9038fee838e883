#!/bin/bash
# usage: gpu_run_multi.sh N   (N GPUs): scene-parallel bench with and without the fused all-reduce kernel
N=$1
mkdir -p gpurun_out
for f in 1 0; do
SPG_FUSED_ALLREDUCE=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 8 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/r2_bench_${N}gpu_f$f.json 2> gpurun_out/r2_bench_${N}gpu_f$f.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_${N}gpu_f$f.json').read().strip().splitlines()[-1])
    print("${N}gpu fused_allreduce=$f", d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['clocks'])
except Exception as ex:
    print("bench failed", ex); print(open('gpurun_out/r2_bench_${N}gpu_f$f.err').read()[-1500:])
PY
done
