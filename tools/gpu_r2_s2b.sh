#!/bin/bash
# round 2, session 2, one GPU: native backtrace of the installable-tool crash; launch list of the link-list build at 8 M
mkdir -p gpurun_out
AQUA_SEGV_BACKTRACE=1 timeout 300 python -m pytest tests/test_installable.py -x -q -m gpu > gpurun_out/r2_installable_bt.log 2>&1; echo "rc=$?"
grep -v "^  File\|site-packages" gpurun_out/r2_installable_bt.log | head -60
AQUA_SEGV_BACKTRACE=1 timeout 600 python -m pytest tests/test_gpu_checkpoint.py -x -q -m gpu > gpurun_out/r2_checkpoint.log 2>&1; echo "checkpoint rc=$?"
tail -30 gpurun_out/r2_checkpoint.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sort_|minmax|heads" -c 40 --csv --log-file gpurun_out/r2_launches_ll_8M.csv python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 2 --warm 1 --only linklist_only > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_ll_8M.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:60], r[-1], r[-2])
PY
