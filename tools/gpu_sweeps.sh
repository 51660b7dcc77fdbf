#!/bin/bash
# quick loop: kernel parity tests + per-kernel timing of the sweeps
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 600 python tools/kbench.py --n 1000000 --reps 5 --only neighs,interactions,shepard,lapp,full,lapp_corr,mls,bie_interactions,bie_p_boundary > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log
