#!/bin/bash
# usage: gpu_ab2.sh TAG "engine ring;engine ring;..." ONLY [ncu "engine ring"]
mkdir -p gpurun_out
TAG=$1; CFGS=$2; ONLY=${3:-interactions,shepard,fused_fluid,lapp_corr}; NCU=$4
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
IFS=';' read -ra ARR <<< "$CFGS"
for cfg in "${ARR[@]}"; do
  set -- $cfg
  echo "== engine $1 ring $2" | tee -a gpurun_out/kbench_$TAG.log
  AQC_SWEEP_ENGINE=$1 AQC_SWEEP_RING=$2 timeout 600 python tools/kbench.py --n 1000000 --reps 5 --only $ONLY 2>&1 | grep -v '"case"' | tee -a gpurun_out/kbench_$TAG.log
done
if [ -n "$NCU" ]; then
  set -- $NCU
  AQC_SWEEP_ENGINE=$1 AQC_SWEEP_RING=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 2 -c 1 -o gpurun_out/prof_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 2 --only interactions > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
fi
