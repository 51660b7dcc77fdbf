#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mpi.py tests/test_gpu_kernels.py -x -q -m gpu -k "kernels_match or mask_cache or fused" > gpurun_out/r2_pytest_rlists_s2f.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/r2_pytest_rlists_s2f.log
