#!/bin/bash
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 600 python -m pytest tests/test_gpu_devloop.py -x -q -m gpu > gpurun_out/r2_pytest_lanes.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/r2_pytest_lanes.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --particles ${N:-1000000} --steps 20 --warmup 3 --cpu-n 3000 --cpu-steps 1 \
     > gpurun_out/r2_lanes_$name.json 2> gpurun_out/r2_lanes_$name.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_lanes_$name.json").read().strip().splitlines()[-1])
    print("$name", round(l["ms_per_step"],3), round(l["e2e"]["ms_per_step"],3), l["config"]["device_loops"]["tools_on_second_stream"])
except Exception as e:
    print("$name no line:", e)
PY
}
if [ $rc -eq 0 ]; then
run rows1_1M AQUA_LANE_ROWS=1
run rows0_1M AQUA_LANE_ROWS=0
N=100000 run rows1_100k AQUA_LANE_ROWS=1
N=100000 run rows0_100k AQUA_LANE_ROWS=0
run rows1_prio_hi AQC_LANE1_PRIORITY=-5
fi
