#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ecc_ab.py > gpurun_out/r2_ecc_ab.log 2>&1; cat gpurun_out/r2_ecc_ab.log
python -m pytest tests/test_main_unchanged.py -m gpu -q --timeout 900 > gpurun_out/r2_pytest7.log 2>&1; tail -15 gpurun_out/r2_pytest7.log
