"""Run a resolved case for a few steps on the GPU and print per-step scalars and
per-tool host time.   python tools/run_case.py --n 100000 --steps 5"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import casegen, host

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--maxiter", type=int, default=-1)
a = ap.parse_args()
host.set_log_level(2)
ov = {}
if a.maxiter >= 0:
    ov["iter_midpoint_max"] = a.maxiter
t0 = time.time()
sim, c = casegen.spheric2(a.n, overrides=ov)
print("loaded N=%d in %.2fs, %d tools" % (c["N"], time.time() - t0, len(sim.tools())), flush=True)
for s in range(a.steps):
    t1 = time.time()
    l0 = sim.launch_count()
    sim.step(1); sim.sync()
    print(json.dumps(dict(step=s, ms=round(1e3 * (time.time() - t1), 2), launches=sim.launch_count() - l0,
                          t=float(sim.scalar("t")), dt=float(sim.scalar("dt")),
                          iters=int(sim.scalar("iter_midpoint", np.uint32)),
                          res=float(sim.scalar("Residual_midpoint")),
                          max_neighs=int(sim.scalar("max_neighs", np.uint32)),
                          n_cells=[int(x) for x in sim.scalar("n_cells", np.uint32, 4)])), flush=True)
tt = sorted(sim.tool_times(), key=lambda x: -x[2])[:12]
for name, n, ms in tt:
    print("  %-40s used %5d  host %.2f ms" % (name, n, ms))
r = sim.download("r", unsorted=True)
print("r finite:", bool(np.isfinite(r).all()), "rho range", sim.download("rho").min(), sim.download("rho").max())
