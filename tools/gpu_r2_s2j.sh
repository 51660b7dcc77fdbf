#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_presets.py -x -q -m gpu -k "symmetry or standing_wave" > gpurun_out/r2_pytest_symmetry.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/r2_pytest_symmetry.log
