#!/bin/bash
# A/B of the sweep engines: parity tests on the default engine, then kernel timings
mkdir -p gpurun_out
TAG=${1:-ab}
ONLY=${2:-interactions,shepard,fused_fluid,lapp_corr,mls,bie_interactions,bie_p_boundary}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
for cfg in "2 0" "3 0" "3 6" "3 8" "3 12" "3 16" "3 24"; do
  set -- $cfg
  echo "== engine $1 ring $2" | tee -a gpurun_out/kbench_$TAG.log
  AQC_SWEEP_ENGINE=$1 AQC_SWEEP_RING=$2 timeout 600 python tools/kbench.py --n 1000000 --reps 5 --only $ONLY 2>&1 | grep -v '"case"' | tee -a gpurun_out/kbench_$TAG.log
done
