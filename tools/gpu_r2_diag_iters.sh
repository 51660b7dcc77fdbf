#!/bin/bash
# why does the 4-slab run need 6 inner iterations?  4 ranks without plan reuse, 2 ranks as in round 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpi.py -x -q -m gpu -k 'n_gpus or dead_peer' > gpurun_out/r2_pytest_mpi_4gpu_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_4gpu_b.log
tail -12 gpurun_out/r2_pytest_mpi_4gpu_b.log
run() { # tag nranks extra-env
  env $3 AQ_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 3 --warmup 5 > gpurun_out/r2_diag_$1.json 2> gpurun_out/r2_diag_$1.err
  echo "$1 rc=$?"; grep "warm-up" gpurun_out/r2_diag_$1.err; python - <<PY
import json
for l in open("gpurun_out/r2_diag_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$1", d["ms_per_step"], d["config"]["mean_inner_iterations"], d["config"]["one_gpu_same_pipeline"])
PY
}
run n4_noplans 4 AQC_MPI_PLANS=0
run n2 2 AQC_MPI_PLANS=1
