#!/bin/bash
# the whole 1-GPU suite + smoke + the default bench line at HEAD
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_1gpu_s3a.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_1gpu_s3a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_s3a.json 2> gpurun_out/r2_bench_1gpu_s3a.err
echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench_1gpu_s3a.json
