#!/bin/bash
# two-lane device loops: parity, then before / after at 1 M and 100 k particles
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 600 python -m pytest tests/test_gpu_devloop.py -x -q -m gpu > gpurun_out/r2_pytest_lanes.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -30 gpurun_out/r2_pytest_lanes.log
if [ $rc -eq 0 ]; then
  for n in 1000000 100000; do
    for ln in 0 1; do
      AQUA_DEVICE_LANES=$ln timeout 300 python bench.py --particles $n --steps 20 --warmup 3 --cpu-n 3000 --cpu-steps 1 \
        > gpurun_out/r2_bench_lanes_${n}_l${ln}.json 2> gpurun_out/r2_bench_lanes_${n}_l${ln}.err
      echo "bench n=$n lanes=$ln rc=$?"; grep "second stream" gpurun_out/r2_bench_lanes_${n}_l${ln}.err | head -2
      python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_bench_lanes_${n}_l${ln}.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("ms_per_step", "value", "gpu_launches")}, l["e2e"]["ms_per_step"], l["config"].get("device_loops"))
except Exception as e:
    print("no line:", e)
PY
    done
  done
fi
