#!/bin/bash
# ncu --set full of the fused fluid sweep, launch list of one bench step, lattice kernel timings
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep3 -s 2 -c 1 -o gpurun_out/prof_fused_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 2 --only fused_fluid > gpurun_out/ncu_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --cpu-n 3000 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 5 --only linklist,sort_stage1+2,predictor,eos,interactions,shepard,rates,corrector,timestep,reduce_min > gpurun_out/kbench_lattice8M_$TAG.log 2>&1
cat gpurun_out/kbench_lattice8M_$TAG.log
