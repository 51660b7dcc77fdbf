#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-b1}
for c in 0 1; do
  AQC_PAIR_CACHE=$c timeout 900 python bench.py --cpu-n 3000 > gpurun_out/bench_${TAG}_cache$c.log 2>&1; echo "bench rc=$?"
  tail -1 gpurun_out/bench_${TAG}_cache$c.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config'].get('pair_mask_cache'), d['roofline']['ms_per_launch'], d['gpu_launches'])"
done
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
