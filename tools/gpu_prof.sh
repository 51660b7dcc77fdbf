#!/bin/bash
# launch list of one bench step + per-kernel timings
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python tools/kbench.py --n 1000000 --reps 5 --only interactions,shepard,fused_fluid,lapp_corr,mls,bie_interactions,bie_p_boundary,bie_elastic_bounce,bie_pst,neighs,linklist,sort_stage1+2 > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --cpu-n 3000 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu rc=$?"
