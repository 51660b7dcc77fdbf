#!/bin/bash
# BASELINE config 5 weak scaling: gpurun --gpus N -- 'bash tools/gpu_lattice_multi.sh N [n_side] [hfac]'
# (one n_side^3 block per GPU, z slabs, 76-tool pipeline over NCCL) + the 2-GPU parity test.
mkdir -p gpurun_out
N=${1:-2}; NS=${2:-200}; HF=${3:-2}
timeout 900 python -m pytest tests/test_gpu_presets.py -q -k "two_gpus" > gpurun_out/pytest_lattice_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_lattice_${N}gpu.log
timeout 600 python tools/bench_lattice.py $NS $HF 10 > gpurun_out/bench_lattice_1gpu_${NS}_hfac$HF.log 2>&1; tail -1 gpurun_out/bench_lattice_1gpu_${NS}_hfac$HF.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/bench_lattice.py $NS $HF 10 > gpurun_out/bench_lattice_${N}gpu_${NS}_hfac$HF.log 2>&1; tail -1 gpurun_out/bench_lattice_${N}gpu_${NS}_hfac$HF.log
