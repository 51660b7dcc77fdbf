#!/bin/bash
# round 2, session 2, one GPU: the new single-kernel-per-digit sort first (fail fast), the whole -m gpu suite,
# the bench as the driver runs it, link-list / permutation alone at 1.2 M and at 8 M lattice particles
mkdir -p gpurun_out
TAG=${1:-s2a}
timeout 600 python -m pytest tests/test_gpu_linklist.py tests/test_gpu_mpi.py -x -q -m gpu -k "linklist or radix or kernels_match" > gpurun_out/r2_pytest_sort_$TAG.log 2>&1; echo "sort tests rc=$?"
tail -5 gpurun_out/r2_pytest_sort_$TAG.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu_$TAG.log
tail -6 gpurun_out/r2_pytest_1gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu_$TAG.json 2> gpurun_out/r2_bench_1gpu_$TAG.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_1gpu_$TAG.json; tail -3 gpurun_out/r2_bench_1gpu_$TAG.err
timeout 600 python tools/kbench.py --n 1000000 --reps 10 --only linklist,linklist_only,sort_stage1+2 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_1M_$TAG.jsonl
timeout 600 python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 10 --only linklist,linklist_only,sort_stage1+2 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_8M_$TAG.jsonl
cat gpurun_out/r2_kbench_ll_1M_$TAG.jsonl gpurun_out/r2_kbench_ll_8M_$TAG.jsonl
