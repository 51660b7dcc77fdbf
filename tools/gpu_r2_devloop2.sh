#!/bin/bash
mkdir -p gpurun_out
AQC_LOOP_DEBUG=1 timeout 300 python bench.py --particles 1000000 --steps 3 --warmup 3 --pre-steps 5 --cpu-n 3000 --cpu-steps 1 \
   > gpurun_out/r2_devloop_debug.json 2> gpurun_out/r2_devloop_debug.err
echo rc=$?; grep "aqc_loop" gpurun_out/r2_devloop_debug.err | tail -12
timeout 300 python -m pytest tests/test_gpu_devloop.py -x -q -m gpu 2>&1 | tail -3
for n in 100000 1000000; do
  AQUA_DEVICE_LOOPS=1 timeout 300 python bench.py --particles $n --steps 20 --warmup 3 --cpu-n 3000 --cpu-steps 1 \
    > gpurun_out/r2_bench_devloop_${n}_dl1b.json 2> gpurun_out/r2_bench_devloop_${n}_dl1b.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/r2_bench_devloop_${n}_dl1b.json").read().strip().splitlines()[-1])
print($n, {k: l[k] for k in ("ms_per_step", "value", "gpu_launches")}, l["e2e"]["ms_per_step"], l["config"].get("device_loops"))
PY
done
