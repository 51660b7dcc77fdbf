#!/bin/bash
# A/B of sweep-engine builds: parity on the in-tree build, then kernel timings of every variant
mkdir -p gpurun_out
TAG=${1:-v}
ONLY=${ONLY:-shepard,fused_fluid,lapp_corr,mls,build+shepard}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_r2.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2_pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_$TAG.log
  tail -5 gpurun_out/r2_pytest_$TAG.log
fi
cp aquagpusph_b200/libaquacuda.so /tmp/base.so
for v in base $VARIANTS; do
  if [ $v = base ]; then cp /tmp/base.so aquagpusph_b200/libaquacuda.so; else cp build/variants/$v/libaquacuda.so aquagpusph_b200/libaquacuda.so; fi
  echo "== variant $v" | tee -a gpurun_out/r2_kbench_$TAG.log
  timeout 600 python tools/kbench.py --n 1000000 --reps 5 --cache 1 --only $ONLY 2>&1 | grep -v '"case"' | tee -a gpurun_out/r2_kbench_$TAG.log
done
cp /tmp/base.so aquagpusph_b200/libaquacuda.so
