#!/bin/bash
# one GPU: config 4 after the fusion across p_boundary, the whole -m gpu suite, the bench as the driver runs it
mkdir -p gpurun_out
timeout 600 python tools/bench2d.py 2000000 10 tld > gpurun_out/r2_cfg_c4_tld_2M_fused4.log 2>&1; grep '^{' gpurun_out/r2_cfg_c4_tld_2M_fused4.log | tail -1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1gpu_s2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu_s2g.log
tail -6 gpurun_out/r2_pytest_1gpu_s2g.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_s2g.json 2> gpurun_out/r2_bench_1gpu_s2g.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2_bench_1gpu_s2g.json; tail -3 gpurun_out/r2_bench_1gpu_s2g.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_s2g.json 2>&1; tail -c 600 gpurun_out/r2_bench_reference_s2g.json
