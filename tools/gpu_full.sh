#!/bin/bash
# round check: all GPU tests, smoke, default bench, kernel timings
mkdir -p gpurun_out
TAG=${1:-full}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
tail -3 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$TAG.log
tail -3 gpurun_out/bench_$TAG.log
timeout 600 python tools/kbench.py --n 1000000 --reps 5 > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log
