#!/bin/bash
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 300 python -m pytest tests/test_installable.py -x -q -s -m gpu > gpurun_out/r2_installable_bt.log 2>&1; echo "installable rc=$?"
grep -v "^  File\|site-packages\|frozen" gpurun_out/r2_installable_bt.log | head -60
