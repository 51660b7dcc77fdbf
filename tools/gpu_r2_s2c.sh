#!/bin/bash
# round 2, session 2, one GPU: installable-tool crash (native backtrace), vectorised prepare / heads kernels
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 300 python -m pytest tests/test_installable.py -x -q -m gpu > gpurun_out/r2_installable_bt.log 2>&1; echo "installable rc=$?"
grep -v "^  File\|site-packages\|frozen" gpurun_out/r2_installable_bt.log | head -50
timeout 600 python -m pytest tests/test_gpu_linklist.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2_pytest_sort_s2c.log 2>&1; echo "sort tests rc=$?"
tail -4 gpurun_out/r2_pytest_sort_s2c.log
timeout 600 python tools/kbench.py --n 1000000 --reps 10 --only linklist_only 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_1M_s2c.jsonl
timeout 600 python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 10 --only linklist_only 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_8M_s2c.jsonl
cat gpurun_out/r2_kbench_ll_1M_s2c.jsonl gpurun_out/r2_kbench_ll_8M_s2c.jsonl
