#!/bin/bash
# multi-GPU bench: usage tools/gpu_multi.sh <ngpus> [n per gpu] [steps]
NG=$1; N=${2:-200000}; ST=${3:-3}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps $ST --warmup 3 --particles $N --cpu-n 5000 > gpurun_out/bench_g${NG}_n${N}.log 2>&1; echo "rc=$?" >> gpurun_out/bench_g${NG}_n${N}.log
tail -12 gpurun_out/bench_g${NG}_n${N}.log
