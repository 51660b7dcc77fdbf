#!/bin/bash
# round 2, one GPU: single-GPU lines of the other BASELINE configs (1: 2-D dam break with BI, 4: 2-D tuned liquid
# damper at 2 M, 5: lattice 200^3 at hfac 1.3 / 2 / 3), each with the per-tool profile of one run and an
# ncu --set full capture of its dominant kernel; launch overhead of the 3-D dam break at 100 k particles
mkdir -p gpurun_out
run() { # tag, command...
  TAG=$1; shift
  timeout 600 "$@" > gpurun_out/r2_cfg_$TAG.log 2>&1; echo "$TAG rc=$?"; grep '^{' gpurun_out/r2_cfg_$TAG.log | tail -1
}
run c1_dambreak2d_50k python tools/bench2d.py 50000 50
run c4_tld_2M python tools/bench2d.py 2000000 10 tld
for HF in 1.3 2 3; do run c5_lattice_200_hfac$HF python tools/bench_lattice.py 200 $HF 5; done
AQUA_PROFILE_SYNC=1 timeout 600 python tools/bench2d.py 2000000 5 tld > gpurun_out/r2_cfg_c4_tld_2M_tools.log 2>&1
AQUA_PROFILE_SYNC=1 timeout 600 python tools/bench2d.py 50000 20 > gpurun_out/r2_cfg_c1_dambreak2d_tools.log 2>&1
tail -9 gpurun_out/r2_cfg_c4_tld_2M_tools.log; tail -9 gpurun_out/r2_cfg_c1_dambreak2d_tools.log
# launch lists (which kernel dominates) + one full capture each
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c4_tld_2M.csv python tools/bench2d.py 2000000 2 tld > /dev/null 2>&1; echo "launches c4 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c5_lattice_200.csv python tools/bench_lattice.py 200 2 2 > /dev/null 2>&1; echo "launches c5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_c1_dambreak2d.csv python tools/bench2d.py 50000 10 > /dev/null 2>&1; echo "launches c1 rc=$?"
python - <<'PY'
import csv, collections
for f in ("c4_tld_2M", "c5_lattice_200", "c1_dambreak2d"):
    tot = collections.Counter(); cnt = collections.Counter()
    try:
        for r in csv.reader(open("gpurun_out/r2_launches_%s.csv" % f)):
            if len(r) > 5 and r[0].isdigit():
                k = r[4].split("(")[0][-70:]
                tot[k] += float(r[-1]); cnt[k] += 1
    except Exception as e:
        print(f, e); continue
    s = sum(tot.values())
    print(f, "sum of kernel time %.3f ms over %d launches" % (s / 1e6, sum(cnt.values())))
    for k, v in tot.most_common(6):
        print("   %6.2f %%  x%-5d %s" % (100 * v / s, cnt[k], k))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep" -s 40 -c 1 -o gpurun_out/r2_prof_c5_lattice python tools/bench_lattice.py 200 2 2 > gpurun_out/r2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep" -s 60 -c 1 -o gpurun_out/r2_prof_c4_tld python tools/bench2d.py 2000000 2 tld > gpurun_out/r2_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
# launch overhead at 100 k particles (what a CUDA graph of the inner iteration could win): step time against the
# sum of the kernel durations of the same steps
timeout 600 python bench.py --particles 100000 --steps 20 --warmup 3 --cpu-n 3000 --cpu-steps 1 > gpurun_out/r2_bench_100k.json 2> gpurun_out/r2_bench_100k.err; echo "bench 100k rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench_100k.csv python bench.py --particles 100000 --steps 5 --warmup 3 --pre-steps 2 --cpu-n 3000 --cpu-steps 1 > gpurun_out/r2_launches_bench_100k.log 2>&1; echo "launch list 100k rc=$?"
ls -la gpurun_out | tail -15
