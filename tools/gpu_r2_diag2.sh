#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpi.py -x -q -m gpu -k 'n_gpus or dead_peer' > gpurun_out/r2_pytest_mpi_4gpu_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_4gpu_c.log
tail -12 gpurun_out/r2_pytest_mpi_4gpu_c.log
if grep -q "failed" gpurun_out/r2_pytest_mpi_4gpu_c.log; then
  AQC_MPI_PLANS=0 timeout 900 python -m pytest tests/test_gpu_mpi.py -x -q -m gpu -k 'n_gpus' > gpurun_out/r2_pytest_mpi_4gpu_c_noplans.log 2>&1
  tail -12 gpurun_out/r2_pytest_mpi_4gpu_c_noplans.log
fi
run() { # tag nranks extra-env
  env $3 AQ_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 3 --warmup 5 > gpurun_out/r2_diag_$1.json 2> gpurun_out/r2_diag_$1.err
  echo "$1 rc=$?"; grep "warm-up" gpurun_out/r2_diag_$1.err; python - <<PY
import json
for l in open("gpurun_out/r2_diag_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$1", d["ms_per_step"], d["config"]["mean_inner_iterations"], d["e2e"]["mean_inner_iterations"], d["e2e"]["ms_per_step"], d["config"]["one_gpu_same_pipeline"])
PY
}
run n4_plans_verify 4 "AQC_MPI_PLANS=1 AQC_MPI_VERIFY=1"
