#!/bin/bash
# round 2, 2 GPUs: why one particle of the delta-SPH slab run has no single-GPU twin (tools/diag_dsph.py), and
# the per-tool profile of the delta-SPH slab pipeline
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python tools/diag_dsph.py $N 50000 4 > gpurun_out/r2_diag_dsph_default.log 2>&1
grep -v "^NCCL" gpurun_out/r2_diag_dsph_default.log | head -60
for v in AQC_MPI_PLANS=0 AQC_PAIR_CACHE=0 AQC_REMOTE_NEAR=0 AQUA_NO_FUSION=1 AQC_SWEEP_ENGINE=2; do
  env $v AQ_DIAG_ONLY=4 timeout 300 python tools/diag_dsph.py $N 50000 4 > gpurun_out/r2_diag_dsph_$v.log 2>&1
  grep "steps 4" -A3 gpurun_out/r2_diag_dsph_$v.log | head -8
done
AQUA_PROFILE_SYNC=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/prof_slabs.py 1000000 dsph > gpurun_out/r2_prof_slabs_${N}gpu_dsph.log 2>&1
grep "^rank 0" gpurun_out/r2_prof_slabs_${N}gpu_dsph.log | head -40
