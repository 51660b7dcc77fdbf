#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel list for kbench --only> [n]
mkdir -p gpurun_out
TAG=$1; ONLY=$2; N=${3:-1000000}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 2 -c 1 -o gpurun_out/prof_$TAG python tools/kbench.py --n $N --reps 1 --warm 2 --only $ONLY > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_$TAG.log
tail -4 gpurun_out/ncu_$TAG.log
ls -la gpurun_out
