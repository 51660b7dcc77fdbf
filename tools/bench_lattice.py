"""ms/step of BASELINE config 5: uniform lattice through the pipelines of cases_xml/src/lattice*_3d.
One GPU: n_side^3 fluid particles (200^3 = 8e6 is the per-GPU size BASELINE.json names), 36 tools.
    python tools/bench_lattice.py [n_side] [hfac] [steps]
N GPUs (weak scaling: one n_side^3 block per GPU stacked along z, 76-tool pipeline with the reference's
migration + halo presets over NCCL, time = max over ranks):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_lattice.py [n_side] [hfac] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import _lib, casegen, host

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hfac = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
host.set_log_level(3)
dist = None
if world > 1:
    import datetime
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                            timeout=datetime.timedelta(seconds=180))
    uid = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim, case = casegen.lattice_slab(n_side, rank, world, hfac, device=local_rank, unique_id=uid[0],
                                     nz_local=n_side)
    n_particles = case["n_fluid"] * world
else:
    sim, case = casegen.lattice(n_side, hfac, device=local_rank)
    n_particles = case["N"]
ctx = _lib.Context.borrow(sim.cuda_ctx(), 3)


def barrier():
    sim.sync()
    if dist is not None:
        dist.barrier()


for _ in range(3):
    sim.step(1)
barrier()
e0, e1 = ctx.event(), ctx.event()
l0 = sim.launch_count()
ctx.record(e0)
for _ in range(steps):
    sim.step(1)
ctx.record(e1)
barrier()
ms = ctx.elapsed_ms(e0, e1) / steps
if dist is not None:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    print(json.dumps({"case": "lattice_3d" if world == 1 else "lattice_mpi_3d", "n_gpus": world,
                      "N": n_particles, "hfac": hfac, "ms_per_step": round(ms, 4),
                      "particle_steps_per_s": round(n_particles / ms * 1e3),
                      "launches_per_step_rank0": (sim.launch_count() - l0) // steps, "scaling": "weak"}))
    if os.environ.get("AQUA_PROFILE_SYNC"):
        for name, k, t in sorted(sim.tool_times(), key=lambda x: -x[2])[:40]:
            print("  %-40s x%-4d %.3f ms/step" % (name, k, t / (steps + 3)))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
sim.close()
