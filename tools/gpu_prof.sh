#!/bin/bash
# round-2 profiling pass (one GPU): launch list of an eager training step + ncu --set full of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_train_step.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-roofline --no-cpu-baseline --no-parity --no-e2e > gpurun_out/prof_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 1 -c 5 -f -o gpurun_out/r2_tc_gemm2 \
    python tools/tc_one.py 120576 256 128 > gpurun_out/prof_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 2 -c 2 -f -o gpurun_out/r2_tc_gemm2_1m \
    python tools/tc_one.py 1286144 256 128 >> gpurun_out/prof_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_dw -c 1 -f -o gpurun_out/r2_tc_dw \
    python tools/tc_one.py 120576 256 128 >> gpurun_out/prof_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pointnet_fused -s 6 -c 2 -f -o gpurun_out/r2_fused \
    python tools/fused_one.py 18248 > gpurun_out/prof_fused.log 2>&1
ncu --set full --clock-control none -k regex:ecc_ -c 12 -f -o gpurun_out/r2_ecc python tools/profile_ecc.py > gpurun_out/prof_ecc.log 2>&1
python tools/fused_one.py 18248; python tools/fused_one.py 942
ls -la gpurun_out/*.ncu-rep
