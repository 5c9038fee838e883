#!/bin/bash
# 2 GPUs: the multi-GPU tests, then the scene-parallel bench with the fused all-reduce kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 > gpurun_out/multi2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi2_pytest.log
tail -4 gpurun_out/multi2_pytest.log
bash tools/gpu_run_multi.sh 2
