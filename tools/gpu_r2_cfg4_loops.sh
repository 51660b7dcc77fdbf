#!/bin/bash
# BASELINE config 4 (2-D tuned liquid damper, 2 M particles, five passes per step): host loop / device loop / two lanes
mkdir -p gpurun_out
for v in "AQUA_DEVICE_LOOPS=0" "AQUA_DEVICE_LANES=0" "AQUA_DEVICE_LANES=1"; do
  env $v timeout 200 python tools/bench2d.py 2000000 10 tld 2>/dev/null | grep '^{' | tail -1 | sed "s/^/$v /"
done | tee gpurun_out/r2_cfg4_loops.txt
