#!/bin/bash
# round 2, one GPU: the parity suite, then the v4 (neighbour list) engine against the round-1 forms
# of the pair cache, kernel by kernel, the FP32 peak, the bench and one ncu capture of the fused sweep
mkdir -p gpurun_out
TAG=${1:-a}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu_$TAG.log
  tail -8 gpurun_out/r2_pytest_1gpu_$TAG.log
fi
ONLY=interactions,shepard,fused_fluid,lapp_corr,mls,build+shepard
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== cache $1 lists $2" | tee -a gpurun_out/r2_kbench_$TAG.log
  AQC_PAIR_LISTS=$2 timeout 600 python tools/kbench.py --n 1000000 --reps 5 --cache $1 --only $ONLY 2>&1 | grep -v '"case"' | tee -a gpurun_out/r2_kbench_$TAG.log
done
python -c "
from aquagpusph_b200 import _lib
c=_lib.Context(0,dims=3,h=1.0); print('fp32 peak TFLOP/s (FFMA, FFMA2):', c.fp32_peak()); c.close()" 2>&1 | tee gpurun_out/r2_fp32_peak_$TAG.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_$TAG.json 2> gpurun_out/r2_bench_1gpu_$TAG.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2_bench_1gpu_$TAG.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep4 -s 2 -c 1 -o gpurun_out/r2_prof_fused_$TAG python tools/kbench.py --n 1000000 --reps 1 --warm 2 --cache 1 --only fused_fluid > gpurun_out/r2_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -12
