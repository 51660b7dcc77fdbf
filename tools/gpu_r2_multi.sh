#!/bin/bash
# round 2, multi-GPU check: the N-rank tests, then the bench at N ranks (driver's configuration)
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_mpi.py -x -q -m gpu > gpurun_out/r2_pytest_mpi_${N}gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_${N}gpu.log
tail -15 gpurun_out/r2_pytest_mpi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps ${2:-5} --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_${N}gpu.json
tail -5 gpurun_out/r2_bench_${N}gpu.err
