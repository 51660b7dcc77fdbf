#!/bin/bash
# round 2, session 2, N GPUs: every multi-rank test, the bench as the driver launches it, the per-tool profile of the
# delta-SPH slab pipeline, BASELINE config 5 (lattice, z slabs) weak scaling at 200^3 per GPU
N=${1:-2}
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
timeout 1500 python -m pytest tests/test_gpu_mpi.py tests/test_gpu_presets.py -x -q -m gpu -k "${KSEL:-gpus or dead_peer or slabs or mpi_plane or random_masks}" > gpurun_out/r2_pytest_mpi_${N}gpu_s2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_${N}gpu_s2.log
tail -8 gpurun_out/r2_pytest_mpi_${N}gpu_s2.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 10 --warmup 3 $BENCHFLAGS > gpurun_out/r2_bench_${N}gpu_s2.json 2> gpurun_out/r2_bench_${N}gpu_s2.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_${N}gpu_s2.err
python - <<PY
import json
for l in open("gpurun_out/r2_bench_${N}gpu_s2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench N=$N", round(d["ms_per_step"],3), d["value"], d["config"]["mean_inner_iterations"], d["config"]["one_gpu_same_pipeline"], d["roofline"].get("stages"))
PY
if [ "${PROF:-1}" = "1" ]; then
AQUA_PROFILE_SYNC=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/prof_slabs.py 1000000 dsph > gpurun_out/r2_prof_slabs_${N}gpu_dsph_s2.log 2>&1
grep "^rank 0" gpurun_out/r2_prof_slabs_${N}gpu_dsph_s2.log | grep "mpi\|ms/step over\|link" | head -24
fi
if [ "${LATTICE:-1}" = "1" ]; then
NS=${NS:-200}; HF=2
if [ "${LATTICE_ONE:-0}" = "1" ]; then timeout 600 python tools/bench_lattice.py $NS $HF 10 > gpurun_out/r2_bench_lattice_1gpu_${NS}.log 2>&1; tail -1 gpurun_out/r2_bench_lattice_1gpu_${NS}.log; fi
for HF in ${HFS:-2}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/bench_lattice.py $NS $HF 10 > gpurun_out/r2_bench_lattice_${N}gpu_${NS}_hfac$HF.log 2>&1; tail -1 gpurun_out/r2_bench_lattice_${N}gpu_${NS}_hfac$HF.log
done
fi
