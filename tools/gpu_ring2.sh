#!/bin/bash
# usage: gpu_ring2.sh TAG "variant:ring2 ..." ONLY   (mask-reading sweeps: tiles per round x ring rounds)
mkdir -p gpurun_out
TAG=$1; VARS=$2; ONLY=${3:-fused_fluid,shepard,lapp_corr,mls}
cp aquagpusph_b200/libaquacuda.so /tmp/libaquacuda_default.so
for vr in $VARS; do
  v=${vr%%:*}; ring=${vr##*:}
  cp build/variants/$v/libaquacuda.so aquagpusph_b200/libaquacuda.so
  echo "== variant $v ring2 $ring" | tee -a gpurun_out/kbench_$TAG.log
  AQC_SWEEP_RING2=$ring timeout 600 python tools/kbench.py --n 1000000 --reps 5 --cache 1 --only $ONLY 2>&1 | grep -v '"case"\|pairs_cache' | cut -c1-${CUT:-60} | tee -a gpurun_out/kbench_$TAG.log
done
cp /tmp/libaquacuda_default.so aquagpusph_b200/libaquacuda.so
