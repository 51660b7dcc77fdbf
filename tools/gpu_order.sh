#!/bin/bash
mkdir -p gpurun_out
TAG=$1; ONLY=${2:-fused_fluid,shepard,lapp_corr,mls}
for so in 0 -1 2 3 4; do
  echo "== suborder $so" | tee -a gpurun_out/kbench_$TAG.log
  timeout 600 python tools/kbench.py --n 1000000 --reps 5 --suborder $so --only $ONLY 2>&1 | grep -v '"case"' | cut -c1-60 | tee -a gpurun_out/kbench_$TAG.log
done
