#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
AQUA_PROFILE_SYNC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/bench_lattice.py 200 2 5 > gpurun_out/r2_prof_lattice_${N}gpu_200.log 2>&1; grep -v "^\*\|OMP\|^$" gpurun_out/r2_prof_lattice_${N}gpu_200.log | tail -45
