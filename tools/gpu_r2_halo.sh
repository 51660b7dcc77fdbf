#!/bin/bash
# round 2, N GPUs: the multi-rank tests, the halo path variants (remote sweep engine) and the per-tool
# profile of the slab pipeline
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mpi.py tests/test_gpu_presets.py -x -q -m gpu -k "${KSEL:-gpus or dead_peer or slabs or mpi_plane or random_masks}" > gpurun_out/r2_pytest_mpi_${N}gpu_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_${N}gpu_h.log
tail -6 gpurun_out/r2_pytest_mpi_${N}gpu_h.log
run() { # tag extra-env
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_halo_${N}gpu_$1.json 2> gpurun_out/r2_halo_${N}gpu_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
for l in open("gpurun_out/r2_halo_${N}gpu_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$1", round(d["ms_per_step"],3), d["config"]["mean_inner_iterations"], d["config"]["one_gpu_same_pipeline"])
PY
}
run dsph AQC_REMOTE_NEAR=1
tail -c 600 gpurun_out/r2_halo_${N}gpu_dsph.err
