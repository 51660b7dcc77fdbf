#!/bin/bash
# one GPU: config 4 with the four fluid sweeps fused across the BIe block, the pipeline tests it touches, and the
# ncu --set full captures of the dominant kernels of configs 4 and 5
mkdir -p gpurun_out
timeout 600 python tools/bench2d.py 2000000 10 tld > gpurun_out/r2_cfg_c4_tld_2M_fused4.log 2>&1; grep '^{' gpurun_out/r2_cfg_c4_tld_2M_fused4.log | tail -1
timeout 600 python tools/bench2d.py 50000 50 > gpurun_out/r2_cfg_c1_dambreak2d_50k_b.log 2>&1; grep '^{' gpurun_out/r2_cfg_c1_dambreak2d_50k_b.log | tail -1
timeout 1200 python -m pytest tests/test_gpu_presets.py tests/test_gpu_bi.py tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2_pytest_fusion_s2h.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_fusion_s2h.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep4_kernel.*PFusedFluid" -s 4 -c 1 -o gpurun_out/r2_prof_c4_tld python tools/bench2d.py 2000000 2 tld > gpurun_out/r2_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep3_kernel.*PFusedFluid" -s 0 -c 1 -o gpurun_out/r2_prof_c5_lattice python tools/bench_lattice.py 200 2 8 > gpurun_out/r2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la gpurun_out/*.ncu-rep
