#!/bin/bash
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 900 python -m pytest tests/test_gpu_scripts.py -x -q -m gpu > gpurun_out/r2_pytest_scripts.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/r2_pytest_scripts.log
