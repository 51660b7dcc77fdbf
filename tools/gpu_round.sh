#!/bin/bash
# full check of the round: GPU tests, smoke, bench (cache on/off), launch list of one step, kernel timings
mkdir -p gpurun_out
TAG=${1:-rnd}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$TAG.log
tail -2 gpurun_out/bench_$TAG.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --cpu-n 3000 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python tools/kbench.py --n 1000000 --reps 5 --cache 1 > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log | cut -c1-110
