#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2_pytest14.log 2>&1; tail -15 gpurun_out/r2_pytest14.log
for f in 1 0; do
SPG_FUSED_ALLREDUCE=$f timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-roofline --no-cpu-baseline --no-parity > gpurun_out/r2_bench_2gpu_f$f.json 2> gpurun_out/r2_bench_2gpu_f$f.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_2gpu_f$f.json').read().strip().splitlines()[-1])
    print("2gpu fused_allreduce=$f", d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
except Exception as ex:
    print("bench failed", ex); print(open('gpurun_out/r2_bench_2gpu_f$f.err').read()[-1500:])
PY
done
