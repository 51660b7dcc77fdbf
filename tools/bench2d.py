"""ms/step of the 2-D SPHERIC test 5 dam break (BI boundary integrals, 57-tool pipeline,
BASELINE config 1) at a given size.   python tools/bench2d.py [n_reservoir] [steps]"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import _lib, cases, casegen, host
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
host.set_log_level(3)
case = cases.spheric5_dam_break_2d(n, 3.0)
sim = casegen.load("spheric5_dambreak_2d", case, (case["N"],))
ctx = _lib.Context.borrow(sim.cuda_ctx(), 2)
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
print(json.dumps({"case": "spheric5_dambreak_2d", "N": case["N"], "ms_per_step": round(ms, 4),
                  "particle_steps_per_s": round(case["N"] / ms * 1e3), "launches_per_step": (sim.launch_count() - l0) // steps,
                  "engine": os.environ.get("AQC_SWEEP_ENGINE", "3")}))
tt = sorted(sim.tool_times(), key=lambda x: -x[2])[:8]
if os.environ.get("AQUA_PROFILE_SYNC"):
    for name, k, t in tt:
        print("  %-40s x%-4d %.3f ms/step" % (name, k, t / (steps + 3)))
