#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sweep4_kernel.*PFusedFluid" -s 4 -c 1 -o gpurun_out/r2_prof_c4_tld python tools/bench2d.py 2000000 2 tld > gpurun_out/r2_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sweep3_kernel.*PFusedFluid" -s 0 -c 1 -o gpurun_out/r2_prof_c5_lattice python tools/bench_lattice.py 200 2 8 > gpurun_out/r2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_ncu_c4.log gpurun_out/r2_ncu_c5.log
