#!/bin/bash
mkdir -p gpurun_out
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
run base AQUA_DEVICE_LANES=1
run prio_hi AQC_LANE1_PRIORITY=-5
run gain1 AQUA_LANE_GAIN=1
run gain8 AQUA_LANE_GAIN=8
run gain20 AQUA_LANE_GAIN=20
run sweep3 AQUA_LANE_SWEEP_COST=3
run sweep30 AQUA_LANE_SWEEP_COST=30
