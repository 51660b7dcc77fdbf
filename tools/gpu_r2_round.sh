#!/bin/bash
# round 2, one GPU: the whole -m gpu suite, the bench as the driver runs it, the launch list of one bench
# command and ncu --set full of the dominant kernel (neighbour-list reader) and of the list builder
mkdir -p gpurun_out
TAG=${1:-f}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu_$TAG.log
tail -6 gpurun_out/r2_pytest_1gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu_$TAG.json 2> gpurun_out/r2_bench_1gpu_$TAG.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench_1gpu_$TAG.json; tail -3 gpurun_out/r2_bench_1gpu_$TAG.err
timeout 900 python tools/kbench.py --n 1000000 --reps 5 --cache 1 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_1M_$TAG.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_1M_$TAG.csv python bench.py --steps 2 --warmup 3 --pre-steps 2 --cpu-n 3000 --cpu-steps 1 > gpurun_out/r2_launches_bench_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep4 -s 2 -c 1 -o gpurun_out/r2_prof_fused_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 2 --cache 1 --only fused_fluid > gpurun_out/r2_ncu_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:PMaskBuild -s 1 -c 1 -o gpurun_out/r2_prof_builder_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 1 --cache 1 --only build+shepard > gpurun_out/r2_ncu_builder_$TAG.log 2>&1; echo "ncu builder rc=$?"
ls -la gpurun_out | tail -12
