"""Per-tool time (AQUA_PROFILE_SYNC=1) of the slab pipeline on N GPUs, rank 0 and the last rank.
    AQUA_PROFILE_SYNC=1 python -m torch.distributed.run --nproc-per-node 2 ... tools/prof_slabs.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from aquagpusph_b200 import casegen, host
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
uid = [None]
if world > 1:
    import datetime
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr), timeout=datetime.timedelta(seconds=120))
    uid = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
host.set_log_level(3)
dsph = len(sys.argv) > 2 and sys.argv[2] == "dsph"
sim, case = casegen.spheric2_slab(n * world, rank, world, device=lr, unique_id=uid[0], delta_sph=dsph)
for _ in range(3):
    sim.step(1)
sim.sync()
t0 = {name: (k, ms) for name, k, ms in sim.tool_times()}
steps = 5
for _ in range(steps):
    sim.step(1)
sim.sync()
rows = []
for name, k, ms in sim.tool_times():
    k0, ms0 = t0.get(name, (0, 0.0))
    if k > k0:
        rows.append((ms - ms0, k - k0, name))
rows.sort(reverse=True)
if world > 1:
    dist.barrier()
if rank in (0, world - 1):
    tot = sum(r[0] for r in rows)
    sys.stdout.write("rank %d: %.2f ms/step over %d tools\n" % (rank, tot / steps, len(rows)))
    for ms, k, name in rows[:44]:
        sys.stdout.write("rank %d  %-40s x%-3d %8.3f ms/step\n" % (rank, name, k // steps, ms / steps))
    sys.stdout.flush()
if world > 1:
    dist.destroy_process_group()
