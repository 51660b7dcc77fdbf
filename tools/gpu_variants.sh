#!/bin/bash
# usage: gpu_variants.sh TAG "variant[:ring] variant[:ring] ..." [ONLY] [test-variant]
# A/B of build/variants/<variant>/libaquacuda.so (tools/build_variant.py) on the kernel bench
mkdir -p gpurun_out
TAG=$1; VARS=$2; ONLY=${3:-fused_fluid,shepard,lapp_corr,mls}; TESTV=$4
cp aquagpusph_b200/libaquacuda.so /tmp/libaquacuda_default.so
for vr in $VARS; do
  v=${vr%%:*}; ring=""; [[ "$vr" == *:* ]] && ring=${vr##*:}
  cp build/variants/$v/libaquacuda.so aquagpusph_b200/libaquacuda.so
  echo "== variant $v ring ${ring:-default}" | tee -a gpurun_out/kbench_$TAG.log
  AQC_SWEEP_RING=$ring timeout 600 python tools/kbench.py --n 1000000 --reps 5 $KB_ARGS --only $ONLY 2>&1 | grep -v '"case"' | cut -c1-${CUT:-60} | tee -a gpurun_out/kbench_$TAG.log
done
if [ -n "$TESTV" ]; then
  cp build/variants/$TESTV/libaquacuda.so aquagpusph_b200/libaquacuda.so
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
  tail -5 gpurun_out/pytest_$TAG.log
fi
cp /tmp/libaquacuda_default.so aquagpusph_b200/libaquacuda.so
