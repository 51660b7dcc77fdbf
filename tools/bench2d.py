"""ms/step of the 2-D cases at a given size.
    python tools/bench2d.py [n] [steps]            SPHERIC test 5 dam break (BI boundary integrals, 57-tool
                                                   pipeline, BASELINE config 1; n = reservoir particles)
    python tools/bench2d.py [n] [steps] tld        SPHERIC test 9 tuned liquid damper (BIe + forces + energy +
                                                   motion presets, hfac 4, BASELINE config 4; n = fluid
                                                   particles, 2000000 in BASELINE.json; prescribed roll of
                                                   casegen.prescribed_roll instead of the python tools)
    python tools/bench2d.py [n] [steps] cavity     SPHERIC test 3 lid-driven cavity (BI + BINoSlip, hfac 4;
                                                   n = fluid particles, 40000 as shipped)"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import _lib, cases, casegen, host
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
host.set_log_level(3)
which = sys.argv[3] if len(sys.argv) > 3 else "dambreak"
if which == "tld":
    name = "spheric9_tld_2d"
    sim, case = casegen.spheric9_tld(n, 4.0)
elif which == "cavity":
    name = "spheric3_liddriven_2d"
    sim, case = casegen.spheric3_lid_driven(int(round(n ** 0.5)), 4.0)
else:
    name = "spheric5_dambreak_2d"
    case = cases.spheric5_dam_break_2d(n, 3.0)
    sim = casegen.load(name, case, (case["N"],))
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
print(json.dumps({"case": name, "N": case["N"], "ms_per_step": round(ms, 4),
                  "particle_steps_per_s": round(case["N"] / ms * 1e3), "launches_per_step": (sim.launch_count() - l0) // steps,
                  "mean_inner_iterations": (float(sim.scalar("iter_midpoint", np.uint32)) if which == "tld" else None),
                  "engine": os.environ.get("AQC_SWEEP_ENGINE", "3")}))
tt = sorted(sim.tool_times(), key=lambda x: -x[2])[:8]
if os.environ.get("AQUA_PROFILE_SYNC"):
    for name, k, t in tt:
        print("  %-40s x%-4d %.3f ms/step" % (name, k, t / (steps + 3)))
