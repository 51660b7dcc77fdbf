import sys, time, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
from aquagpusph_b200 import _lib, casegen, host
import bench
host.set_log_level(3)
sim, case = casegen.spheric2(1000000, device=0)
actx = _lib.Context.borrow(sim.cuda_ctx(), 3)
for _ in range(3): sim.step(1)
sim.sync()
fields = ["r", "u", "dudt", "rho", "drhodt", "m", "imove"]; outs = ["r", "u", "rho", "p"]
hin = {}
for k in fields:
    cur = sim.download(k, np.int32 if k == "imove" else np.float32)
    hin[k] = bench.pinned(actx, cur.shape, cur.dtype); hin[k][...] = cur
hout = {k: bench.pinned(actx, hin[k].shape if k in hin else (case["N"],), np.float32) for k in outs}
for it in range(3):
    t0 = time.perf_counter()
    for k in fields: sim.upload(k, hin[k])
    t1 = time.perf_counter()
    sim.step(1)
    t2 = time.perf_counter()
    sim.sync()
    t3 = time.perf_counter()
    for k in outs: sim.download(k, hout[k].dtype, out=hout[k])
    t4 = time.perf_counter()
    dt = float(sim.scalar("dt"))
    t5 = time.perf_counter()
    print("upload %.2f ms  step(host) %.2f  sync %.2f  download %.2f  scalar %.2f" % tuple(1e3*x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4)), flush=True)
