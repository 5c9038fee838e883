#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 2>&1 | grep -i "fused\|FAIL\|ALL" > gpurun_out/r2_tc_check2.log; cat gpurun_out/r2_tc_check2.log
timeout 600 python tools/tc_bench.py > gpurun_out/r2_tc_bench2.log 2>&1; cat gpurun_out/r2_tc_bench2.log
timeout 600 python tools/grad_diag.py > gpurun_out/r2_grad_diag.log 2>&1; cat gpurun_out/r2_grad_diag.log
