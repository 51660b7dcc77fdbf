#!/bin/bash
# large configurations on one GPU: 16 M particle dam break (config 3 at N = 1) and the 8 M lattice kernels (config 5)
mkdir -p gpurun_out
TAG=${1:-big}
timeout 1200 python bench.py --particles 16000000 --steps 2 --warmup 3 --cpu-n 3000 > gpurun_out/bench_16M_$TAG.log 2>&1; echo "bench16M rc=$?"
tail -1 gpurun_out/bench_16M_$TAG.log | cut -c1-900
nvidia-smi --query-gpu=memory.used --format=csv
timeout 600 python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 3 --cache 1 --only linklist,sort_stage1+2,predictor,eos,interactions,shepard,fused_fluid,rates,corrector,timestep,reduce_min > gpurun_out/kbench_lattice8M_$TAG.log 2>&1
cat gpurun_out/kbench_lattice8M_$TAG.log | cut -c1-120
