"""GPU: the multi-device path.  (a) one GPU: every cfd/MPI.cl and cfd/MPI/planes.cl
kernel through the Kernel-tool C-ABI against the reference's own scripts
(oracle/_ref); (b) two GPUs (skipped with fewer): the reference's tests/2D/MPI_plane
case on two processes -- mpi-sync over NCCL, one process per GPU -- against the
two-rank CPU oracle and against the single-GPU run."""
import os
import socket
import sys

import numpy as np
import pytest

import cases
import mpi_common
from aquagpusph_b200 import _lib
from oracle import ref

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dims", [2, 3])
def test_mpi_kernels_match_reference_scripts(oracle, dims):
    case = cases.dam_break(dims, 14 if dims == 3 else 60, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    n_radix = 1
    while n_radix < ((N + 1023) // 1024) * 1024:
        n_radix *= 2
    rng = np.random.default_rng(11)
    ll = oracle.linklist(case["r"], dims, 2.0, case["h"])
    inv = ll["inv_perm"]
    s = {k: oracle.scatter(case[k], inv) for k in ("imove", "iset", "r", "u", "dudt", "rho", "drhodt", "m")}
    nrecv = N // 3
    host = dict(s)
    host["p"] = rng.normal(size=N).astype(np.float32) * 100
    for k, shape in (("mpi_r", (n_radix, V)), ("mpi_u", (n_radix, V)), ("mpi_dudt", (n_radix, V)),
                     ("mpi_r_in", (n_radix, V)), ("mpi_u_in", (n_radix, V))):
        host[k] = np.zeros(shape, np.float32)
    for k in ("mpi_rho", "mpi_drhodt", "mpi_m", "mpi_p", "mpi_rho_in", "mpi_m_in"):
        host[k] = np.ones(n_radix, np.float32)
    for k in ("mpi_iset", "mpi_iset_in", "mpi_local_mask", "mpi_neigh_mask", "mpi_icell", "mpi_ihoc",
              "mpi_id_sorted", "mpi_id_unsorted"):
        host[k] = np.zeros(n_radix, np.uint32)
    host.update(icell=ll["icell"], ihoc=ll["ihoc"], n_cells=ll["ncells"], N=N, n_radix=n_radix,
                mpi_rank=1, nbuffer=int((s["imove"] == -255).sum()) + nrecv, cs=case["cs"], p0=0.0,
                refd=case["refd"], domain_max=case["domain_max"], r_max=ll["rmax"],
                mpi_plane_proc=0)
    pr = np.zeros(V, np.float32)
    pr[0] = float(np.median(s["r"][s["imove"] == 1][:, 0]))
    pn = np.zeros(V, np.float32)
    pn[0] = -1.0
    host.update(mpi_plane_r=pr, mpi_plane_n=pn)
    host["grad_p"] = rng.normal(size=(N, V)).astype(np.float32)
    host["lap_u"] = rng.normal(size=(N, V)).astype(np.float32)
    host["div_u"] = rng.normal(size=N).astype(np.float32)
    host["shepard"] = rng.uniform(0.5, 1, N).astype(np.float32)

    R = ref.Ref(dims, case["h"])
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    dev = {k: (ctx.array(v) if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] in (N, n_radix)
               else v) for k, v in host.items()}
    dev["refd"] = ctx.array(case["refd"])
    hostv = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in host.items()}

    def both(script, entry, n=None, exact=True, outs=()):
        R.run(script, entry, n if n is not None else n_radix, hostv)
        ctx.launch(script, entry, dev, n=n if n is not None else n_radix)
        for k in outs:
            a, b = hostv[k], dev[k].get()
            if exact:
                assert np.array_equal(a, b, equal_nan=True), (script, entry, k)
            else:
                sc = np.abs(a).max()
                assert np.abs(a.astype(np.float64) - b).max() <= 5e-6 * sc + 1e-30, (script, entry, k)

    for name in ("mpi_local_mask", "mpi_neigh_mask"):
        hostv[name][...] = 1
        dev[name].set(hostv[name])
    both("cfd/MPI/planes.cl", "local_mask", outs=("mpi_local_mask",))
    both("cfd/MPI/planes.cl", "neigh_mask", outs=("mpi_neigh_mask",))
    assert 0 < (hostv["mpi_neigh_mask"][:N] == 0).sum() < N
    both("cfd/MPI.cl", "copy", outs=("mpi_iset", "mpi_r", "mpi_u", "mpi_dudt", "mpi_rho", "mpi_drhodt", "mpi_m"))
    # pretend the first nrecv rows arrived from process 0
    m = np.ones(n_radix, np.uint32)
    m[:nrecv] = 0
    for name in ("mpi_local_mask", "mpi_neigh_mask"):
        hostv[name][...] = m
        dev[name].set(m)
    both("cfd/MPI.cl", "append", outs=("iset", "r", "u", "dudt", "rho", "drhodt", "m", "imove"))
    both("cfd/MPI.cl", "remove", outs=("imove", "r", "u", "dudt", "m"))
    both("cfd/MPI.cl", "backup_r", n=n_radix, outs=("mpi_r_in",))
    for k in ("mpi_iset", "mpi_u", "mpi_rho", "mpi_m"):
        hostv[k + "_in"][...] = hostv[k]
        dev[k + "_in"].set(hostv[k])
    # remote link-list on the LOCAL grid (recompute_grid="false")
    icell, perm, invp = (ctx.empty(n_radix, np.uint32) for _ in range(3))
    _, _, nc, ihoc = ctx.linklist(dev["mpi_r_in"], 2.0, case["h"], icell, None, perm, invp,
                                  rmin=ll["rmin"], rmax=ll["rmax"], recompute=False)
    rl = oracle.linklist(hostv["mpi_r_in"], dims, 2.0, case["h"], ll["rmin"], ll["rmax"], recompute=False)
    assert np.array_equal(icell.get(), rl["icell"]) and np.array_equal(invp.get(), rl["inv_perm"])
    assert np.array_equal(ihoc.get()[:nc[3]], rl["ihoc"][:nc[3]])
    hostv.update(mpi_icell=rl["icell"], mpi_ihoc=rl["ihoc"], mpi_id_sorted=rl["inv_perm"])
    dev.update(mpi_icell=icell, mpi_ihoc=ihoc, mpi_id_sorted=invp)
    both("cfd/MPI.cl", "sort", outs=("mpi_iset", "mpi_r", "mpi_u", "mpi_rho", "mpi_m"))
    both("cfd/MPI.cl", "eos", outs=("mpi_p",))
    pre = {k: dev[k].get() for k in ("shepard", "grad_p", "lap_u", "div_u")}
    both("cfd/MPI.cl", "gamma", n=N, exact=False, outs=("shepard",))
    both("cfd/MPI.cl", "interactions", n=N, exact=False, outs=("grad_p", "lap_u", "div_u"))
    # the fused remote sweep (interactions + gamma in one pass, what the pipeline launches)
    # must add the same terms as its members
    sep = {k: dev[k].get() for k in pre}
    for k, v in pre.items():
        dev[k].set(v)
    ctx.launch_fused([("cfd/MPI.cl", "interactions"), ("cfd/MPI.cl", "gamma")], dev)
    for k in pre:
        a, b = sep[k].astype(np.float64), dev[k].get().astype(np.float64)
        assert np.abs(a - pre[k]).max() > 0, k
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max(), ("fused remote sweep", k)
    # ---- aqua/MPIdeltaSPH.cl (ours: the remote delta-SPH / MLS terms the reference's MPI preset
    # lacks) against its plain-C statement in the oracle, on the same halo list
    from oracle import oracle as O
    D = O.make_defs(dims, case["h"])
    rl = O.make_ll(hostv["mpi_icell"], hostv["mpi_ihoc"], hostv["n_cells"], N)
    extra = {"mls": rng.normal(size=(N, V * V)).astype(np.float32),
             "lap_p": rng.normal(size=N).astype(np.float32),
             "lap_p_corr": rng.normal(size=(N, V)).astype(np.float32),
             "mpi_lap_p_corr": rng.normal(size=(n_radix, V)).astype(np.float32),
             "mpi_lap_p_corr_in": rng.normal(size=(n_radix, V)).astype(np.float32)}
    for k, v in extra.items():
        hostv[k] = v.copy()
        dev[k] = ctx.array(v)
    hostv["mls_imove"] = dev["mls_imove"] = 1

    def ours(entry, call, outs, exact=False):
        call()
        ctx.launch("aqua/MPIdeltaSPH.cl", entry, dev, n=N)
        for k in outs:
            a, b = hostv[k].astype(np.float64), dev[k].get().astype(np.float64)
            if exact:
                assert np.array_equal(a, b), (entry, k)
            else:
                assert np.abs(a - extra[k]).max() > 0, (entry, k, "no remote term")
                assert np.abs(a - b).max() <= 5e-6 * np.abs(a).max(), (entry, k)

    H = hostv
    ours("mls", lambda: O.pcall("mpi_mls", N, D, rl, H["icell"], H["imove"], H["r"], H["mpi_r"], H["mpi_rho"],
                                H["mpi_m"], H["mls"], 1), ("mls",))
    pre = {k: dev[k].get() for k in ("shepard", "grad_p", "lap_u", "div_u", "lap_p", "lap_p_corr")}
    ours("full_lapp", lambda: O.pcall("mpi_dsph_full_lapp", N, D, rl, H["icell"], H["imove"], H["r"], H["p"],
                                      H["mpi_r"], H["mpi_rho"], H["mpi_m"], H["mpi_p"], H["lap_p_corr"],
                                      H["lap_p"]), ("lap_p_corr", "lap_p"))
    # the pipeline launches interactions + gamma + full_lapp as ONE remote sweep
    ctx.launch("cfd/MPI.cl", "gamma", dev, n=N)
    ctx.launch("cfd/MPI.cl", "interactions", dev, n=N)
    sep = {k: dev[k].get() for k in pre}
    for k, v in pre.items():
        dev[k].set(v)
    ctx.launch_fused([("cfd/MPI.cl", "interactions"), ("cfd/MPI.cl", "gamma"),
                      ("aqua/MPIdeltaSPH.cl", "full_lapp")], dev)
    for k in pre:
        a, b = sep[k].astype(np.float64), dev[k].get().astype(np.float64)
        assert np.abs(a - pre[k]).max() > 0, k
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max(), ("fused remote sweep with delta-SPH", k)
    extra["lap_p"] = dev["lap_p"].get()
    H["lap_p"][...] = extra["lap_p"]
    H["lap_p_corr"][...] = dev["lap_p_corr"].get()
    ours("lapp_corr", lambda: O.pcall("mpi_dsph_lapp_corr", N, D, rl, H["icell"], H["imove"], H["r"],
                                      H["lap_p_corr"], H["mpi_r"], H["mpi_rho"], H["mpi_m"], H["mpi_lap_p_corr"],
                                      H["lap_p"]), ("lap_p",))

    def copy_g():
        H["mpi_lap_p_corr"][:N] = H["lap_p_corr"][:N]

    def sort_g():
        H["mpi_lap_p_corr"][H["mpi_id_sorted"][:N]] = H["mpi_lap_p_corr_in"][:N]
    ours("copy_g", copy_g, ("mpi_lap_p_corr",), exact=True)
    ours("sort_g", sort_g, ("mpi_lap_p_corr",), exact=True)
    # ---- the remote sweeps on neighbour lists of their own (the second pair cache, what the host
    # turns on): one build serves every remote sweep while the halo list and the local geometry
    # stand; same pairs as the filtering sweeps above, summed in another order
    outs = ("shepard", "grad_p", "lap_u", "div_u", "lap_p", "lap_p_corr", "mls")
    start = {k: dev[k].get() for k in outs}

    def remote_sweeps():
        ctx.launch("aqua/MPIdeltaSPH.cl", "mls", dev, n=N)
        ctx.launch_fused([("cfd/MPI.cl", "interactions"), ("cfd/MPI.cl", "gamma"),
                          ("aqua/MPIdeltaSPH.cl", "full_lapp")], dev)
        ctx.launch("aqua/MPIdeltaSPH.cl", "lapp_corr", dev, n=N)
        ctx.launch("cfd/MPI.cl", "interactions", dev, n=N)
        return {k: dev[k].get().astype(np.float64) for k in outs}
    filtered = remote_sweeps()
    for k, v in start.items():
        dev[k].set(v)
    ctx.pairs_cache(True)
    listed = remote_sweeps()
    st = ctx.pairs_cache_stats(remote=True)
    if os.environ.get("AQC_REMOTE_LISTS", "1") != "0" and os.environ.get("AQC_SWEEP_ENGINE", "3") == "3":
        assert st["builds"] in (1, 2) and st["hits"] >= 4, st   # (mls asks for the fluid only, the fused sweep widens it)
    for k in outs:
        assert np.abs(filtered[k] - start[k]).max() > 0, k
        assert np.abs(filtered[k] - listed[k]).max() <= 2e-6 * np.abs(filtered[k]).max(), ("remote lists", k)
    # a new halo list drops them (the link-list build reports the arrays it writes)
    ctx.linklist(dev["mpi_r_in"], 2.0, case["h"], icell, ihoc, perm, invp, rmin=ll["rmin"], rmax=ll["rmax"],
                 recompute=False)
    ctx.launch("cfd/MPI.cl", "interactions", dev, n=N)
    st2 = ctx.pairs_cache_stats(remote=True)
    if st["builds"]:
        assert st2["builds"] == st["builds"] + 1, (st, st2)
    ctx.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpu_rank(rank, port, q, fixed_mask, overrides, steps):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=2)
    from aquagpusph_b200 import cases as cs, casegen, host
    host.set_log_level(3)
    table = np.load(os.path.join(HERE, "golden", "reference_inputs.npz"))["mpi_plane_2D"]
    uid = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    arrays, own = cs.mpi_plane(table, rank)
    import mpi_common as mc
    import tempfile
    d = tempfile.mkdtemp()
    path = os.path.join(d, "mpi.rank%d.xml" % rank)
    open(path, "w").write(mc.mpi_xml(overrides, fixed_mask))
    sim = host.Simulation(path, dims=2, device=rank, mpi_rank=rank, mpi_size=2)
    for k, a in arrays.items():
        sim.upload(k, a)
    sim.comm_init(uid[0])
    sim.step(steps)
    res = {k: sim.download(k, np.int32 if k == "imove" else np.float32, unsorted=True) for k in mc.FIELDS}
    res["own"] = own
    res["launches"] = sim.launch_count()
    q.put((rank, res))
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


def _run_two_gpus(fixed_mask, overrides, steps=1):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_rank, args=(r, port, q, fixed_mask, overrides, steps)) for r in range(2)]
    [p.start() for p in procs]
    return _collect(procs, q, 2, 600)


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_mpi_plane_two_gpus(golden, oracle):
    if not _two_gpus():
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    table = golden["mpi_plane_2D"]
    # (1) the reference's test as shipped: serial == 2 processes to 1e-6, and GPU == oracle
    got = _run_two_gpus(False, None)
    assert all(got[r]["launches"] > 0 for r in range(2))
    mpi_common.check_against_serial(mpi_common.oracle_serial(table), got, 1e-6)
    want = mpi_common.oracle_two_ranks_threads(table)
    for r in range(2):
        for k in mpi_common.FIELDS:
            a, b = np.asarray(want[r][k], np.float64), np.asarray(got[r][k], np.float64)
            assert np.abs(a - b).max() <= 1e-6 * max(np.abs(a).max(), 1e-30), (r, k)
    # (2) rates that count (relax 0) and the halo mask taken after the sort:
    #     two GPUs reproduce the serial run and the two-rank oracle
    ov = {"relax_midpoint": "0.0"}
    got = _run_two_gpus(True, ov)
    serial = mpi_common.oracle_serial(table, 1, ov)
    mpi_common.check_against_serial(serial, got, 2e-5, relative=True)
    want = mpi_common.oracle_two_ranks_threads(table, 1, ov, fixed_mask=True)
    for r in range(2):
        for k in mpi_common.FIELDS:
            a, b = np.asarray(want[r][k], np.float64), np.asarray(got[r][k], np.float64)
            assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-30), (r, k)


def _slab_rank(rank, size, port, q, n_total, steps, kw=None):
    import torch.distributed as dist
    os.environ["AQC_MPI_VERIFY"] = "1"   # every reused mpi-sync plan is checked against its mask
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from aquagpusph_b200 import casegen, host
    host.set_log_level(3)
    uid = [None]
    if size > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=size)
        uid = [host.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
    ov = {"iter_midpoint_max": int(os.environ.get("AQ_DIAG_MAXITER", "2"))}
    kw = dict(kw or {})
    whole = kw.pop("whole", False)
    if "iter_midpoint_max" in kw:
        ov["iter_midpoint_max"] = kw.pop("iter_midpoint_max")
    sim, c = casegen.spheric2_slab(n_total, rank, size, overrides=ov, device=rank, unique_id=uid[0], **kw)
    sim.step(steps)
    nf = c["n_fluid"]
    if whole:
        # after migration a rank's rows are no longer its initial fluid: the test takes every
        # row of set 0 in device order and matches the live fluid ones by position
        n0 = c["n_set0"]
        res = {k: sim.download(k, np.float32)[:n0] for k in ("r", "u", "rho", "dudt")}
        res["imove"] = sim.download("imove", np.int32)[:n0]
    else:
        res = {k: sim.download(k, np.float32, unsorted=True)[:nf] for k in ("r", "u", "rho", "dudt")}
        res["imove"] = sim.download("imove", np.int32, unsorted=True)[:nf]
    res["sync_tools"] = {name: n for name, n, _ in sim.tool_times() if "sync" in name}
    res["plan0"] = _lib.Context.borrow(sim.cuda_ctx(), 3).mpi_sync_stats(0) if size > 1 else None
    res["n_fluid0"] = nf
    res["fluid_index"] = c["fluid_index"]
    res["dt"] = float(sim.scalar("dt"))
    res["slab"] = c["slab"]
    res["h"] = c["h"]
    q.put((rank, res))
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()
    sim.close()


def _collect(procs, q, size, timeout=900):
    """Results of `size` ranks; a rank that dies fails the test at once (and takes the
    others with it) instead of leaving the parent in q.get()."""
    import queue
    import time
    got, t0 = {}, time.time()
    while len(got) < size:
        try:
            rank, res = q.get(timeout=1.0)
            got[rank] = res
        except queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError("rank process(es) failed: exit codes %s after %.0f s"
                                     % ([p.exitcode for p in procs], time.time() - t0))
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return got


def _run_slabs(size, n_total, steps, **kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_rank, args=(r, size, port, q, n_total, steps, kw)) for r in range(size)]
    [p.start() for p in procs]
    return _collect(procs, q, size)


def test_dam_break_slabs_two_gpus_match_one_gpu():
    """BASELINE config 3 shape at test size: the 3-D dam break through the reference's MPI
    example pipeline on 2 GPUs (y slabs, halo + migration over NCCL, global dt, halo
    refreshed every midpoint sub-iteration: casegen.multi_device_fixes) against the same
    pipeline on 1 GPU, two steps of two sub-iterations.  Particles farther than the kernel
    support from the cut never see a remote term: with the order-preserving sweep engine
    (AQC_SWEEP_ENGINE=2) they are bit-identical; the default engine sums a particle's pairs
    in an order that depends on which particles share its CTA, so far from the cut the two
    runs agree to the accumulated fp32 rounding of the sums, like next to the cut (where the
    remote terms are added after the local ones)."""
    if not _two_gpus():
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n_total, steps = 40000, 2
    one = _run_slabs(1, n_total, steps)[0]
    two = _run_slabs(2, n_total, steps)
    assert two[0]["dt"] == two[1]["dt"] == one["dt"], "dt must be global and equal to the 1-GPU value"
    pos1 = {int(g): k for k, g in enumerate(one["fluid_index"])}
    for r in range(2):
        rows = np.array([pos1[int(g)] for g in two[r]["fluid_index"]])
        assert np.array_equal(two[r]["imove"], one["imove"][rows])
        far = np.abs(one["r"][rows][:, 1] - two[0]["slab"][1]) > 6.0 * two[r]["h"]
        assert far.any() and (~far).any()
        for k, tol in (("r", 1e-6), ("u", 5e-5), ("rho", 2e-5), ("dudt", 5e-5)):
            a = one[k][rows].astype(np.float64)
            b = two[r][k].astype(np.float64)
            err = np.abs(a - b).max() / max(np.abs(a).max(), 1e-30)
            assert err <= tol, "rank %d field %s: rel err %.3e" % (r, k, err)
            if os.environ.get("AQC_SWEEP_ENGINE") == "2":
                assert np.array_equal(one[k][rows][far], two[r][k][far]), \
                    "rank %d field %s far from the cut" % (r, k)
            else:
                errf = np.abs(a[far] - b[far]).max() / max(np.abs(a).max(), 1e-30)
                assert errf <= 0.5 * tol, "rank %d field %s far from the cut: rel err %.3e" % (r, k, errf)


# ---------------------------------------------------------------------------------------------
# More than two ranks: an interior rank talks to two peers, and nothing about the counts is
# special (round 1's states were lattices at rest whose halo counts were multiples of 4, which
# hid a misaligned packed send buffer).

def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _sync_rank(rank, size, port, q, seed):
    """aqc_mpi_sync through the C-ABI on random masks: fields of 4, 8 and 16 bytes in that
    order, odd element counts, every rank talks to every other one; then the plan path."""
    import torch.distributed as dist
    os.environ["AQC_MPI_VERIFY"] = "1"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=size)
    uid = [_lib.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = _lib.Context(rank, dims=3, h=1.0)
    ctx.comm_init(rank, size, uid[0])
    n = 4099 + 37 * rank   # (every rank's arrays must hold what it receives: ~n/size per peer)
    rng = np.random.default_rng(seed + rank)
    # ~1/3 of the elements travel; counts per destination are arbitrary (not multiples of 4)
    mask_h = np.where(rng.random(n) < 0.33, rng.integers(0, size, n), rank).astype(np.uint32)
    f4 = (np.arange(n) + 100000 * rank).astype(np.uint32)
    f8 = rng.normal(size=(n, 2)).astype(np.float32)
    f16 = rng.normal(size=(n, 4)).astype(np.float32)
    out = {"mask0": mask_h.copy(), "f4_0": f4.copy(), "f8_0": f8.copy(), "f16_0": f16.copy()}
    mask, a4, a8, a16 = ctx.array(mask_h), ctx.array(f4), ctx.array(f8), ctx.array(f16)
    out["nrecv"] = ctx.mpi_sync(mask, [a4, a8, a16])
    out.update(mask=mask.get(), f4=a4.get(), f8=a8.get(), f16=a16.get())
    # ---- plan: a second call on the same mask content reuses the sort and the counts
    dep = ctx.array(np.zeros(16, np.float32))
    plan = ctx.mpi_sync_plan()
    for it in range(3):
        mask.set(mask_h)
        a4.set(f4 + it)
        a16.set(f16 * (it + 1))
        if it == 2:
            ctx.fill(dep, np.float32(1).tobytes())   # a dependency was written: full call again
        out["nrecv_plan%d" % it] = ctx.mpi_sync(mask, [a4, a16], plan=plan, deps=[dep])
        out["plan%d" % it] = dict(mask=mask.get(), f4=a4.get(), f16=a16.get())
    out["stats"] = ctx.mpi_sync_stats(plan)
    q.put((rank, out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _expected_sync(rank, size, masks, fields):
    """MPISync.cpp:183-232: blocks packed at the front in process order, stable inside."""
    parts = [[] for _ in fields]
    m = []
    for p in range(size):
        if p == rank:
            continue
        sel = np.flatnonzero(masks[p] == rank)
        for k, f in enumerate(fields):
            parts[k].append(f[p][sel])
        m += [p] * len(sel)
    return [np.concatenate(x) for x in parts], np.array(m, np.uint32)


@pytest.mark.parametrize("size", [2, 3, 4])
def test_mpi_sync_random_masks(size):
    if _n_gpus() < size:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (size, size))
    import torch.multiprocessing as mp
    mpx = mp.get_context("spawn")
    q = mpx.Queue()
    port = _free_port()
    procs = [mpx.Process(target=_sync_rank, args=(r, size, port, q, 77)) for r in range(size)]
    [p.start() for p in procs]
    got = _collect(procs, q, size, 300)
    masks = [got[r]["mask0"] for r in range(size)]
    # block sizes that are not multiples of 4 elements (a 4-byte field followed by 16-byte ones)
    assert any(int((masks[p] == r).sum()) % 4 for p in range(size) for r in range(size) if p != r)
    for r in range(size):
        g = got[r]
        (e4, e8, e16), em = _expected_sync(r, size, masks, [[got[p][k] for p in range(size)]
                                                             for k in ("f4_0", "f8_0", "f16_0")])
        k = len(em)
        assert g["nrecv"] == k, (r, k, g["nrecv"])
        assert np.array_equal(g["mask"][:k], em) and np.all(g["mask"][k:] == r)
        assert np.array_equal(g["f4"][:k], e4) and np.array_equal(g["f4"][k:], g["f4_0"][k:])
        assert np.array_equal(g["f8"][:k], e8) and np.array_equal(g["f8"][k:], g["f8_0"][k:])
        assert np.array_equal(g["f16"][:k], e16) and np.array_equal(g["f16"][k:], g["f16_0"][k:])
        for it in range(3):
            (p4, p16), _ = _expected_sync(r, size, masks, [[got[p]["f4_0"] + it for p in range(size)],
                                                           [got[p]["f16_0"] * (it + 1) for p in range(size)]])
            assert g["nrecv_plan%d" % it] == k
            assert np.array_equal(g["plan%d" % it]["mask"][:k], em)
            assert np.array_equal(g["plan%d" % it]["f4"][:k], p4), (r, it)
            assert np.array_equal(g["plan%d" % it]["f16"][:k], p16), (r, it)
        assert g["stats"] == dict(full=2, reused=1), g["stats"]


def _match_rows(one, ranks):
    """Rows of the 1-GPU run (original order: row k is particle fluid_index[k] of the whole dam
    break) that correspond to every live fluid row of the N-GPU run.  Particles migrate, so ids
    are rank-local: the match goes through positions (two particles are never closer than a
    fraction of dr; the runs differ by fp32 rounding)."""
    from scipy.spatial import cKDTree
    fl1 = np.flatnonzero(one["imove"] == 1)
    tree = cKDTree(one["r"][fl1][:, :3].astype(np.float64))
    out = []
    for g in ranks:
        fl = np.flatnonzero(g["imove"] == 1)
        d, j = tree.query(g["r"][fl][:, :3].astype(np.float64))
        out.append((fl, fl1[j], d))
    return fl1, out


@pytest.mark.parametrize("size", [3, 4])
def test_dam_break_slabs_n_gpus_match_one_gpu(size):
    """3 and 4 y-slabs of a JITTERED, moving dam break (halo and migration counts arbitrary;
    particles cross the cuts during the run) against the same pipeline on one GPU: interior
    ranks exchange with two peers, migrating particles land in buffer rows, every reused
    mpi-sync plan is verified against its mask (AQC_MPI_VERIFY)."""
    if _n_gpus() < size:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (size, size))
    # (a slab must be thicker than the two halos it feeds: 1.07 / 4 > 2 x 2 h at this size)
    n_total, steps = 100000, 6
    kw = dict(seed=5, jitter=0.45, uscale=2.0, whole=True, iter_midpoint_max=3)
    one = _run_slabs(1, n_total, steps, **dict(kw, whole=False))[0]
    many = _run_slabs(size, n_total, steps, **kw)
    assert len({many[r]["dt"] for r in range(size)}) == 1, "dt must be global"
    # (against one GPU the velocities dt derives from differ by the rounding of the summation order)
    assert abs(many[0]["dt"] - one["dt"]) <= 1e-6 * one["dt"]
    fl1, matches = _match_rows(one, [many[r] for r in range(size)])
    n_live = sum(len(m[0]) for m in matches)
    assert n_live == len(fl1), "fluid particles lost or duplicated: %d vs %d" % (n_live, len(fl1))
    assert len(np.unique(np.concatenate([m[1] for m in matches]))) == len(fl1)
    h = one["h"]
    # particles crossed the cuts: ranks hold particles that started on another rank
    arrived = [int((~np.isin(one["fluid_index"][matches[r][1]], many[r]["fluid_index"])).sum())
               for r in range(size)]
    assert sum(arrived) > 20, "hardly anything migrated: %s" % arrived
    for r in range(size):
        rows, rows1, d = matches[r]
        assert d.max() < 1e-4 * h, "rank %d: a particle is %.3e away from its 1-GPU twin" % (r, d.max())
        for k, tol in (("r", 2e-6), ("u", 2e-4), ("rho", 5e-5), ("dudt", 5e-4)):
            a = one[k][rows1].astype(np.float64)
            b = many[r][k][rows].astype(np.float64)
            err = np.abs(a - b).max() / max(np.abs(one[k][fl1]).max(), 1e-30)
            assert err <= tol, "rank %d field %s: rel err %.3e" % (r, k, err)
        # the in-loop halo sync sorted and counted once per step and reused that plan afterwards
        calls = many[r]["sync_tools"].get("mpi neighs sync", 0)
        st = many[r]["plan0"]
        assert st["full"] == steps and st["full"] + st["reused"] == calls and st["reused"] >= steps, (st, calls)


def _dead_peer_rank(rank, size, port, q):
    import time
    import torch.distributed as dist
    os.environ["AQC_COMM_TIMEOUT_S"] = "8"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=size)
    uid = [_lib.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = _lib.Context(rank, dims=3, h=1.0)
    ctx.comm_init(rank, size, uid[0])
    n = 1000
    mask = ctx.array(np.full(n, (rank + 1) % size, np.uint32))
    f = ctx.array(np.arange(n, dtype=np.float32))
    assert ctx.mpi_sync(mask, [f]) == n      # the communicator works
    dist.barrier()
    if rank == size - 1:
        os._exit(0)                          # this rank leaves without a word
    mask.set(np.full(n, (rank + 1) % size, np.uint32))
    t0 = time.time()
    try:
        ctx.mpi_sync(mask, [f])
        ctx.sync()
        q.put((rank, ("no error", time.time() - t0, "")))
        q.close()
        q.join_thread()
    except _lib.AquaError as e:
        dt = time.time() - t0
        # and the context stays failed: no later collective may hang either
        try:
            ctx.mpi_sync(mask, [f])
            again = "no error"
        except _lib.AquaError as e2:
            again = str(e2)
        q.put((rank, (str(e), dt, again)))
        q.close()
        q.join_thread()      # (os._exit would drop what the feeder thread has not sent yet)
    os._exit(0)


def test_dead_peer_ends_the_job():
    """A rank that disappears must not leave its peers inside ncclRecv (round 1: three ranks
    spun for 870 s).  Either the bounded wait aborts the communicator and the call fails (and
    every later collective of that context fails at once), or -- when the host is stuck inside
    NCCL itself -- the watchdog thread ends the process with exit code 70; both within the
    time-out (8 s here) plus the grace periods."""
    import queue
    import time
    size = min(_n_gpus(), 3)
    if size < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mpx = mp.get_context("spawn")
    q = mpx.Queue()
    port = _free_port()
    procs = [mpx.Process(target=_dead_peer_rank, args=(r, size, port, q)) for r in range(size)]
    t0 = time.time()
    [p.start() for p in procs]
    [p.join(100) for p in procs]
    took = time.time() - t0
    alive = [p.is_alive() for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert not any(alive), "ranks still alive after %.0f s: %s" % (took, alive)
    got = {}
    while True:
        try:
            r, v = q.get(timeout=0.5)
            got[r] = v
        except queue.Empty:
            break
    for r in range(size - 1):
        if r in got:     # the call failed in line
            msg, dt, again = got[r]
            assert msg != "no error" and dt < 40.0, (r, msg, dt)
            # (a vanished rank can also surface as garbage in the gathered counts: the symmetry
            # check of the process lists then fails the call before any wait expires)
            assert "abort" in msg or "peer" in msg or "NCCL" in msg or "symmetric" in msg, msg
            assert "aborted" in again or "symmetric" in again or "peer" in again, again
        else:            # the watchdog ended the process
            assert procs[r].exitcode == 70, (r, procs[r].exitcode)
    assert took < 90.0, took


def _single_116_rank(q, n_total, steps, kw):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from aquagpusph_b200 import casegen, host
    host.set_log_level(3)
    kw = dict(kw)
    ov = {"iter_midpoint_max": kw.pop("iter_midpoint_max", 3)}
    sim, c = casegen.spheric2(n_total, overrides=ov, device=0, **kw)
    sim.step(steps)
    nf = c["n_fluid"]
    res = {k: sim.download(k, np.float32, unsorted=True)[:nf] for k in ("r", "u", "rho", "dudt")}
    res["imove"] = sim.download("imove", np.int32, unsorted=True)[:nf]
    res.update(dt=float(sim.scalar("dt")), h=c["h"], fluid_index=np.arange(nf), tools=len(sim.tools()))
    q.put((0, res))
    sim.close()


@pytest.mark.parametrize("size", [2, 3])
def test_delta_sph_slabs_match_the_single_gpu_pipeline(size):
    """BASELINE config 2's physics on N GPUs: the slab pipeline with delta-SPH and MLS
    (casegen.slab_delta_sph: remote terms aqua/MPIdeltaSPH.cl, a second halo exchange per
    sub-iteration for the corrected pressure gradient) against the UNCHANGED 116-tool pipeline of
    examples/3D/spheric_testcase2_dambreak on one GPU -- the pipeline bench.py runs at N = 1.
    Jittered, moving particles (they migrate at two ranks, where the cut runs through a lattice
    layer); every live particle within fp32 summation rounding of its single-GPU twin."""
    if _n_gpus() < size:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (size, size))
    import torch.multiprocessing as mp
    n_total, steps = 50000, 4
    kw = dict(seed=5, jitter=0.45, uscale=2.0, iter_midpoint_max=3)
    mpx = mp.get_context("spawn")
    q = mpx.Queue()
    p = mpx.Process(target=_single_116_rank, args=(q, n_total, steps, kw))
    p.start()
    one = _collect([p], q, 1, 600)[0]
    many = _run_slabs(size, n_total, steps, whole=True, delta_sph=True, **kw)
    assert len({many[r]["dt"] for r in range(size)}) == 1
    assert abs(many[0]["dt"] - one["dt"]) <= 1e-6 * one["dt"]
    fl1, matches = _match_rows(one, [many[r] for r in range(size)])
    assert sum(len(m[0]) for m in matches) == len(fl1)
    assert len(np.unique(np.concatenate([m[1] for m in matches]))) == len(fl1)
    for r in range(size):
        rows, rows1, d = matches[r]
        assert d.max() < 1e-4 * one["h"], (r, d.max())
        for k, tol in (("r", 2e-6), ("u", 2e-4), ("rho", 5e-5), ("dudt", 1e-3)):
            a = one[k][rows1].astype(np.float64)
            b = many[r][k][rows].astype(np.float64)
            err = np.abs(a - b).max() / max(np.abs(one[k][fl1]).max(), 1e-30)
            assert err <= tol, "rank %d field %s: rel err %.3e" % (r, k, err)
