#!/bin/bash
# one GPU, quick: the sort with the concurrent look-back (fail fast before the multi-GPU calls)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linklist.py -x -q -m gpu > gpurun_out/r2_pytest_sort_s2e.log 2>&1; echo "sort tests rc=$?"
tail -4 gpurun_out/r2_pytest_sort_s2e.log
timeout 600 python tools/kbench.py --n 1000000 --reps 10 --only linklist_only 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_1M_s2e.jsonl
timeout 600 python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 10 --only linklist_only 2>&1 | grep -v '"case"' > gpurun_out/r2_kbench_ll_8M_s2e.jsonl
cat gpurun_out/r2_kbench_ll_1M_s2e.jsonl gpurun_out/r2_kbench_ll_8M_s2e.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sort_|minmax|heads" -c 40 --csv --log-file gpurun_out/r2_launches_ll_8M_s2e.csv python tools/kbench.py --case lattice --n 8000000 --hfac 2 --reps 2 --warm 1 --only linklist_only > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_ll_8M_s2e.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-5:]: print(r[4][:60], r[-1], r[-2])
PY
