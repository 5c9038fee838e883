#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 1 -c 5 -f -o gpurun_out/r2_tc_gemm2 \
    python tools/tc_one.py 120576 256 128 > gpurun_out/prof_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 2 -c 2 -f -o gpurun_out/r2_tc_gemm2_1m \
    python tools/tc_one.py 1286144 256 128 >> gpurun_out/prof_tc.log 2>&1
tail -3 gpurun_out/prof_tc.log
for spec in "sweep_vv 100000" "vkitti_train 1024"; do set -- $spec
timeout 900 python bench.py --workload $1 --nodes $2 --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench_$1_$2.json 2> gpurun_out/r2_bench_$1_$2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_$1_$2.json').read().strip().splitlines()[-1])
    print("$1 $2", {k:d.get(k) for k in ('ms_per_step','value','gpu_launches','cuda_graph','parity_rel_err')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])
except Exception as ex:
    print("$1 failed", ex); print(open('gpurun_out/r2_bench_$1_$2.err').read()[-1200:])
PY
done
nvidia-smi --query-gpu=memory.used --format=csv
