#!/bin/bash
# device-side loops: parity tests, then the launch + sync overhead per step before / after at 100 k and 1 M particles
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 600 python -m pytest tests/test_gpu_devloop.py -x -q -m gpu > gpurun_out/r2_pytest_devloop.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -40 gpurun_out/r2_pytest_devloop.log
if [ $rc -eq 0 ]; then
  for n in 100000 1000000; do
    for dl in 0 1; do
      AQUA_DEVICE_LOOPS=$dl timeout 300 python bench.py --particles $n --steps 20 --warmup 3 --cpu-n 3000 --cpu-steps 1 \
        > gpurun_out/r2_bench_devloop_${n}_dl${dl}.json 2> gpurun_out/r2_bench_devloop_${n}_dl${dl}.err
      echo "bench n=$n dl=$dl rc=$?"; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_bench_devloop_${n}_dl${dl}.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("ms_per_step", "value", "gpu_launches")}, l["e2e"]["ms_per_step"], l["config"].get("mean_inner_iterations"), l["config"].get("device_loops"))
except Exception as e:
    print("no line:", e)
PY
    done
  done
fi
