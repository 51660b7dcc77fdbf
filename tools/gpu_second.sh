#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/kbench.py --n 1000000 --reps 5 > gpurun_out/kbench_1M.log 2>&1; echo "kbench rc=$?" >> gpurun_out/kbench_1M.log
cat gpurun_out/kbench_1M.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_default.log
tail -3 gpurun_out/bench_default.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:PInteractions -s 2 -c 1 -o gpurun_out/prof_interactions_v1 python tools/kbench.py --n 1000000 --reps 1 --warm 2 --only interactions > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
