"""ms/step of BASELINE config 5 on one GPU: uniform lattice of n_side^3 fluid particles through the
36-tool pipeline of cases_xml/src/lattice_3d (200^3 = 8e6 is the per-GPU size BASELINE.json names).
    python tools/bench_lattice.py [n_side] [hfac] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import _lib, casegen, host

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hfac = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
host.set_log_level(3)
sim, case = casegen.lattice(n_side, hfac)
ctx = _lib.Context.borrow(sim.cuda_ctx(), 3)
for _ in range(3):
    sim.step(1)
sim.sync()
e0, e1 = ctx.event(), ctx.event()
l0 = sim.launch_count()
ctx.record(e0)
for _ in range(steps):
    sim.step(1)
ctx.record(e1)
sim.sync()
ms = ctx.elapsed_ms(e0, e1) / steps
print(json.dumps({"case": "lattice_3d", "N": case["N"], "hfac": hfac, "ms_per_step": round(ms, 4),
                  "particle_steps_per_s": round(case["N"] / ms * 1e3),
                  "launches_per_step": (sim.launch_count() - l0) // steps}))
if os.environ.get("AQUA_PROFILE_SYNC"):
    for name, k, t in sorted(sim.tool_times(), key=lambda x: -x[2])[:8]:
        print("  %-40s x%-4d %.3f ms/step" % (name, k, t / (steps + 3)))
