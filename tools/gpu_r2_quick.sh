#!/bin/bash
# the device-loop tests + one short bench line (a quick check after host-side changes)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_devloop.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-n 3000 --cpu-steps 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['ms_per_step'],3), round(l['e2e']['ms_per_step'],3), l['config']['device_loops'])"
