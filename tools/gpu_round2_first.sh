#!/bin/bash
# First GPU call of the next round: everything written after round 1's GPU budget ran out
# (tests/test_gpu_presets.py: motion / energy kernels, the tuned-liquid-damper pipeline), then the
# usual round check (tools/gpu_round.sh) and BASELINE config 4 / config 1 timings.
mkdir -p gpurun_out
TAG=${1:-r2first}
timeout 600 python -m pytest tests/test_gpu_presets.py -q > gpurun_out/pytest_presets_$TAG.log 2>&1; echo "presets rc=$?" >> gpurun_out/pytest_presets_$TAG.log
tail -30 gpurun_out/pytest_presets_$TAG.log
bash tools/gpu_round.sh $TAG
timeout 600 python tools/bench2d.py 2000000 10 tld > gpurun_out/bench2d_tld_2M_$TAG.log 2>&1; tail -3 gpurun_out/bench2d_tld_2M_$TAG.log
timeout 300 python tools/bench2d.py 50000 50 > gpurun_out/bench2d_dambreak_50k_$TAG.log 2>&1; tail -3 gpurun_out/bench2d_dambreak_50k_$TAG.log
for HF in 1.3 2 3; do
  timeout 600 python tools/bench_lattice.py 200 $HF 5 > gpurun_out/bench_lattice_8M_hfac${HF}_$TAG.log 2>&1; tail -2 gpurun_out/bench_lattice_8M_hfac${HF}_$TAG.log
done
timeout 300 python tools/bench2d.py 1000000 10 cavity > gpurun_out/bench2d_cavity_1M_$TAG.log 2>&1; tail -2 gpurun_out/bench2d_cavity_1M_$TAG.log
