#!/bin/bash
# usage: tools/gpu_ncu2.sh <tag> <kernel regex> <kbench --only list> [extra kbench args]
mkdir -p gpurun_out
TAG=$1; RX=$2; ONLY=$3; shift 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 1 -o gpurun_out/prof_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 3 --only $ONLY "$@" > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_$TAG.log
tail -3 gpurun_out/ncu_$TAG.log
