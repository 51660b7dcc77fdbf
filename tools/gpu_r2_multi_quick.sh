#!/bin/bash
# quick N-GPU check at HEAD: the plane test + the bench as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mpi.py -x -q -m gpu -k "mpi_plane or dead_peer or random_masks" > gpurun_out/r2_pytest_mpi_${N}gpu_quick.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_mpi_${N}gpu_quick.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 5 --warmup 3 --skip-same-pipeline > gpurun_out/r2_bench_${N}gpu_quick.json 2> gpurun_out/r2_bench_${N}gpu_quick.err
echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2_bench_${N}gpu_quick.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench N=$N", round(d["ms_per_step"],3), d["value"], d["config"]["mean_inner_iterations"], d["config"].get("device_loops"))
PY
