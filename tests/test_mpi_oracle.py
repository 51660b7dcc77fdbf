"""CPU: the multi-device path on the oracle.  The reference's own multi-device test
(tests/2D/MPI_plane: one midpoint step of 10000 particles, serial vs split at x = 0
over two processes, every field of every particle equal to 1e-6) is replayed by
the oracle interpreter with two ranks (threads, and separate processes over gloo).
The halo / migration kernels (cfd/MPI.cl, cfd/MPI/planes.cl) are the reference's
own scripts (oracle/_ref)."""
import os
import sys

import numpy as np
import pytest

import mpi_common
from oracle import interp, ref

pytestmark = pytest.mark.skipif(not ref.available() and not ref.build(),
                                reason="oracle/_ref not built (needs /root/reference once)")


def test_mpi_sync_semantics():
    """MPISync.cpp:183-232 on host arrays: stable by destination, packed at the front in
    process order, mask = sender over what arrived and own rank elsewhere."""
    import threading
    size, n = 3, 64
    tr = interp.LocalTransport(size)
    rng = np.random.default_rng(3)
    masks = [rng.integers(0, size, n).astype(np.uint32) for _ in range(size)]
    vals = [np.arange(n, dtype=np.float32) + 1000 * r for r in range(size)]
    out = {}

    def work(rank):
        I = interp.Interpreter.__new__(interp.Interpreter)
        I.rank, I.size, I.transport = rank, size, tr
        I.V = {"mask": masks[rank].copy(), "f": vals[rank].copy()}
        I.eval = lambda e: float(e)
        I.mpi_sync({"mask": "mask", "fields": "f", "processes": ""})
        out[rank] = I.V
    th = [threading.Thread(target=work, args=(k,)) for k in range(size)]
    [t.start() for t in th]
    [t.join() for t in th]
    for rank in range(size):
        want_f, want_m = [], []
        for p in range(size):
            if p == rank:
                continue
            sel = np.flatnonzero(masks[p] == rank)
            want_f += list(vals[p][sel])
            want_m += [p] * len(sel)
        k = len(want_f)
        assert np.array_equal(out[rank]["f"][:k], np.array(want_f, np.float32))
        assert np.array_equal(out[rank]["mask"][:k], np.array(want_m, np.uint32))
        assert np.all(out[rank]["mask"][k:] == rank)
        assert np.array_equal(out[rank]["f"][k:], vals[rank][k:])


def test_mpi_plane_two_ranks_match_serial(golden, oracle):
    table = golden["mpi_plane_2D"]
    serial = mpi_common.oracle_serial(table)
    ranks = mpi_common.oracle_two_ranks_threads(table)
    worst = mpi_common.check_against_serial(serial, ranks, 1e-6)
    # the halo did contribute: without it particles next to x = 0 would differ
    own0 = ranks[0]["own"]
    near = np.abs(np.asarray(serial["r"])[own0][:, 0]) < 0.05
    assert near.any() and np.abs(np.asarray(serial["dudt"])[own0][near]).max() > 0
    assert worst < 1e-6


def test_halo_reproduces_the_serial_run(golden, oracle):
    """With relax_midpoint = 1 (as shipped) the reference's test discards the rates it
    has just computed, so it cannot see the halo.  With relax_midpoint = 0 the new
    rates count; the shipped tool order then fails on the first step because
    mpi_neigh_mask is computed before the sort permutes the particles (see
    casegen.halo_mask_after_sort).  Placing the mask tool after the Sort stage, two
    ranks reproduce the serial run to fp32 summation order."""
    table = golden["mpi_plane_2D"]
    ov = {"relax_midpoint": "0.0"}
    serial = mpi_common.oracle_serial(table, 1, ov)
    stale = mpi_common.oracle_two_ranks_threads(table, 1, ov, fixed_mask=False)
    with pytest.raises(AssertionError):
        mpi_common.check_against_serial(serial, stale, 1e-3, relative=True)
    fixed = mpi_common.oracle_two_ranks_threads(table, 1, ov, fixed_mask=True)
    worst = mpi_common.check_against_serial(serial, fixed, 5e-6, relative=True)
    assert 0 < worst   # remote terms are added after the local ones: not bit-identical


def _gloo_worker(rank, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=2)
    import mpi_common as mc
    from oracle import interp as ip, oracle as O
    O.build()
    table = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "reference_inputs.npz"))["mpi_plane_2D"]
    res = mc.oracle_rank(table, rank, ip.TorchTransport())
    q.put((rank, {k: np.asarray(v) for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_mpi_plane_two_processes_gloo(golden, oracle):
    """world_size = 2 over torch.distributed / gloo: same result as the threaded ranks."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    table = golden["mpi_plane_2D"]
    mpi_common.check_against_serial(mpi_common.oracle_serial(table), got, 1e-6)
    ref_ranks = mpi_common.oracle_two_ranks_threads(table)
    for r in range(2):
        for k in mpi_common.FIELDS:
            assert np.array_equal(got[r][k], ref_ranks[r][k]), (r, k)


def _dam_break_ranks(size, n_total, steps, delta_sph=False, maxiter=3, **kw):
    """The 3-D dam break cut in `size` y slabs through the oracle interpreter (threads), with the
    multi-device additions of casegen.multi_device_fixes (and the delta-SPH / MLS stages of
    casegen.slab_delta_sph); returns the device-order state."""
    import threading
    from aquagpusph_b200 import cases, casegen
    tr = interp.LocalTransport(size)
    out, errs = {}, []

    def work(rank):
        try:
            c = cases.spheric2_dam_break_slab(n_total, 3.0, rank, size, **kw)
            txt = casegen.instantiate("spheric2_dambreak_mpi_3d", c, (c["n_set0"], c["N"] - c["n_set0"]),
                                      {"iter_midpoint_max": maxiter})
            txt = (casegen.slab_fixes_delta_sph(float(c["delta"][0])) if delta_sph
                   else casegen.multi_device_fixes)(txt)
            I = interp.Interpreter(txt, 3, rank=rank, size=size, transport=tr)
            for k in casegen.STATE_FIELDS:
                I.V[k][...] = c[k]
            for _ in range(steps):
                I.step()
            res = {k: I.V[k].copy() for k in ("r", "u", "rho", "dudt", "imove")}
            res.update(fluid_index=c["fluid_index"], n_fluid=c["n_fluid"], dt=float(I.V["dt"]), h=c["h"],
                       r0=I.unsorted("r")[:c["n_fluid"]], u0=I.unsorted("u")[:c["n_fluid"]],
                       unsorted={k: I.unsorted(k)[:c["n_fluid"]] for k in ("r", "u", "rho", "dudt", "drhodt")},
                       n_tools=len(I.tools))
            out[rank] = res
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            tr._barrier.abort()
    th = [threading.Thread(target=work, args=(k,)) for k in range(size)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return out


def test_dam_break_three_slabs_with_migration(oracle):
    """BASELINE config 3's shape with what round 1 never exercised: off-lattice particles that
    move fast enough to cross the cuts, an interior rank with two peers.  Three ranks reproduce
    the one-rank run particle by particle (matched by position: ids are rank-local once a
    particle has migrated), nothing is lost or duplicated, and dozens of particles end on a
    rank they did not start on.  The same scenario runs on GPUs in tests/test_gpu_mpi.py."""
    from scipy.spatial import cKDTree
    from oracle import oracle as O
    O.set_threads(4)
    try:
        kw = dict(seed=5, jitter=0.45, uscale=2.0)
        n_total, steps, size = 24000, 5, 3
        one = _dam_break_ranks(1, n_total, steps, **kw)[0]
        many = _dam_break_ranks(size, n_total, steps, **kw)
    finally:
        O.set_threads(1)
    nf = one["n_fluid"]
    tree = cKDTree(one["r0"][:, :3].astype(np.float64))
    seen, arrived = [], 0
    for r in range(size):
        g = many[r]
        assert g["dt"] == many[0]["dt"] and abs(g["dt"] - one["dt"]) <= 1e-6 * one["dt"]
        fl = np.flatnonzero(g["imove"] == 1)
        d, j = tree.query(g["r"][fl][:, :3].astype(np.float64))
        assert d.max() < 1e-4 * g["h"], (r, d.max())
        seen.append(j)
        arrived += int((~np.isin(one["fluid_index"][j], g["fluid_index"])).sum())
        a, b = one["u0"][j].astype(np.float64), g["u"][fl].astype(np.float64)
        assert np.abs(a - b).max() <= 2e-5 * np.abs(a).max(), r
    seen = np.concatenate(seen)
    assert len(seen) == nf and len(np.unique(seen)) == nf, "particles lost or duplicated"
    assert arrived > 20, arrived


def _one_rank_is_the_116_tool_pipeline_when_particles_reach_the_walls():
    """Jittered, fast particles: some bounce off the walls within a step, so the element radius
    (__DR_FACTOR__, which the single-device example sets and the MPI example does not) must have
    travelled into the slab pipeline with the other definitions."""
    from aquagpusph_b200 import cases, casegen
    kw = dict(seed=5, jitter=0.45, uscale=2.0)
    n_total = 12000
    c = cases.spheric2_dam_break(n_total, 3.0, **kw)
    I = interp.Interpreter(casegen.instantiate("spheric2_dambreak_3d", c, (c["N"] - 8, 8),
                                               {"iter_midpoint_max": 2}), 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    I.step()
    one = _dam_break_ranks(1, n_total, 1, delta_sph=True, maxiter=2, **kw)[0]
    for k in ("r", "u", "rho", "dudt", "drhodt"):
        assert np.array_equal(one["unsorted"][k], I.unsorted(k)[one["fluid_index"]]), k


def test_slab_pipeline_with_delta_sph_is_the_single_device_pipeline(oracle):
    """casegen.slab_delta_sph: the reference's MPI example pipeline extended with the delta-SPH and
    MLS stages of the single-device dam break (remote terms: aqua/MPIdeltaSPH.cl, ours -- the
    reference's MPI preset cannot exchange them).  On ONE rank it is the 116-tool pipeline of
    examples/3D/spheric_testcase2_dambreak bit for bit; on three ranks (an interior rank, cuts off
    the lattice layers so that nothing migrates in two steps) every particle agrees with the
    single-device run to the rounding of a sum whose remote terms come last."""
    from aquagpusph_b200 import cases, casegen
    from oracle import oracle as O
    O.set_threads(4)
    try:
        kw = dict(seed=5, jitter=0.0, uscale=0.1)
        n_total, steps = 24000, 2
        c = cases.spheric2_dam_break(n_total, 3.0, **kw)
        I = interp.Interpreter(casegen.instantiate("spheric2_dambreak_3d", c, (c["N"] - 8, 8),
                                                   {"iter_midpoint_max": 3}), 3)
        for k in casegen.STATE_FIELDS:
            I.V[k][...] = c[k]
        for _ in range(steps):
            I.step()
        serial = {k: I.unsorted(k) for k in ("r", "u", "rho", "dudt", "drhodt")}
        one = _dam_break_ranks(1, n_total, steps, delta_sph=True, **kw)[0]
        three = _dam_break_ranks(3, n_total, steps, delta_sph=True, **kw)
    finally:
        O.set_threads(1)
    assert one["dt"] == float(I.V["dt"])
    for k in serial:
        assert np.array_equal(one["unsorted"][k], serial[k][one["fluid_index"]]), k
    _one_rank_is_the_116_tool_pipeline_when_particles_reach_the_walls()
    for r in range(3):
        g = three[r]
        assert g["dt"] == float(I.V["dt"]) and g["n_tools"] == one["n_tools"]
        for k, tol in (("r", 1e-7), ("u", 2e-5), ("rho", 1e-6), ("dudt", 2e-4), ("drhodt", 2e-4)):
            a = serial[k][g["fluid_index"]].astype(np.float64)
            b = g["unsorted"][k].astype(np.float64)
            assert np.abs(a - b).max() <= tol * np.abs(serial[k]).max(), (r, k)


def _gloo_dsph_worker(rank, port, q, n_total, steps, kw):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=2)
    from aquagpusph_b200 import cases as cs, casegen as cg
    from oracle import interp as ip, oracle as O
    O.build()
    c = cs.spheric2_dam_break_slab(n_total, 3.0, rank, 2, **kw)
    txt = cg.instantiate("spheric2_dambreak_mpi_3d", c, (c["n_set0"], c["N"] - c["n_set0"]), {"iter_midpoint_max": 2})
    txt = cg.slab_fixes_delta_sph(float(c["delta"][0]))(txt)
    I = ip.Interpreter(txt, 3, rank=rank, size=2, transport=ip.TorchTransport())
    for k in cg.STATE_FIELDS:
        I.V[k][...] = c[k]
    for _ in range(steps):
        I.step()
    q.put((rank, {k: I.V[k].copy() for k in ("r", "u", "rho", "dudt", "imove")}))
    dist.barrier()
    dist.destroy_process_group()


def test_delta_sph_slabs_two_processes_gloo(oracle):
    """The delta-SPH slab pipeline (casegen.slab_delta_sph: two halo exchanges per sub-iteration, the pre-loop
    one, migration, both all-reduces) on world_size = 2 over torch.distributed / gloo: bit-identical to the
    same two ranks run as threads of one process -- the transport does not matter, only the protocol."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    kw = dict(seed=5, jitter=0.45, uscale=2.0)
    n_total, steps = 6000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_dsph_worker, args=(r, port, q, n_total, steps, kw)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=600) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = _dam_break_ranks(2, n_total, steps, delta_sph=True, maxiter=2, **kw)
    for r in range(2):
        for k in ("r", "u", "rho", "dudt", "imove"):
            assert np.array_equal(got[r][k], want[r][k]), (r, k)
        assert np.abs(got[r]["dudt"]).max() > 0
