#!/bin/bash
# pair-mask cache: parity tests, kernel timings with and without, bench
mkdir -p gpurun_out
TAG=${1:-c1}; WHAT=${2:-all}
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
for c in 0 1; do
  echo "== cache $c" | tee -a gpurun_out/kbench_$TAG.log
  timeout 600 python tools/kbench.py --n 1000000 --reps 5 --cache $c --only fused_fluid,shepard,lapp_corr,mls,interactions,build+shepard 2>&1 | grep -v '"case"' | cut -c1-200 | tee -a gpurun_out/kbench_$TAG.log
done
if [ "$WHAT" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all_$TAG.log
  tail -5 gpurun_out/pytest_all_$TAG.log
  for c in 0 1; do
    AQC_PAIR_CACHE=$c timeout 900 python bench.py --cpu-n 3000 > gpurun_out/bench_${TAG}_cache$c.log 2>&1; echo "bench rc=$?"
    tail -1 gpurun_out/bench_${TAG}_cache$c.log | cut -c1-400
  done
fi
