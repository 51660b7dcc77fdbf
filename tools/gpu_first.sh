#!/bin/bash
# first GPU contact: tests, smoke, a short bench, a launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --n 200000 --steps 2 --warmup 3 --maxiter 5 --cpu-n 10000 > gpurun_out/bench_200k.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_200k.log
tail -5 gpurun_out/bench_200k.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_200k.csv python bench.py --n 200000 --steps 1 --warmup 3 --maxiter 2 --cpu-n 3000 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_bench.log
tail -3 gpurun_out/ncu_bench.log
