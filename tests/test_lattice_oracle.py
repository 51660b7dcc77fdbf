"""BASELINE config 5 on the CPU: the 36-tool lattice pipeline (cases_xml/src/lattice_3d/Lattice.xml over
the reference's unchanged presets basic + improved_euler + cfd + variableTimeStep) resolves to the
tool list SURVEY 8(d) names, parses identically in the C++ host and the oracle interpreter, names
only registered kernels, and steps in the oracle.  GPU side: tests/test_gpu_presets.py."""
import os
import re
import sys

import numpy as np
import pytest

from aquagpusph_b200 import _lib, casegen, cases, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tool_list(tmp_path):
    from oracle import interp
    c = cases.lattice(10, 2.0)
    txt = casegen.instantiate("lattice_3d", c, (c["N"],))
    p = tmp_path / "lattice.xml"
    p.write_text(txt)
    tools = host.Simulation(str(p), dims=3, parse_only=True).tools()
    assert tools == [(t["name"], t["type"]) for t in interp.Interpreter(txt, 3).tools]
    work = [n for n, t in tools if t not in ("dummy", "copy", "set", "set_scalar", "assert")]
    assert work == ["predictor", "link-list", "sort stage1", "sort stage2", "EOS", "Binormal", "cfd Shepard",
                    "cfd interactions", "cfd sensors", "cfd sensors renormalization", "cfd rates", "corrector",
                    "cfd variable time step", "cfd minimum time step"]
    L = _lib.lib()
    for path, entry in re.findall(r'type="kernel"[^>]*path="[^"]*Scripts/([^"]*)" entry_point="([^"]*)"', txt):
        assert L.aqc_kernel_lookup(path.encode(), entry.encode(), 3) >= 0, (path, entry)
    assert "improved_euler.cl" in txt and "cfd/TimeStep.cl" in txt


@pytest.mark.skipif(not os.path.isdir("/root/reference/resources"),
                    reason="needs the reference tree (build container only)")
def test_committed_template_is_what_the_front_end_resolves(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("resolve_case", os.path.join(ROOT, "tools", "resolve_case.py"))
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    old, rc.OUT = rc.OUT, str(tmp_path)
    try:
        rc.resolve("lattice_3d", *rc.CASES["lattice_3d"])
    finally:
        rc.OUT = old
    assert (tmp_path / "lattice_3d.xml").read_text() == \
        open(os.path.join(ROOT, "aquagpusph_b200", "cases_xml", "lattice_3d.xml")).read()


def test_steps_in_the_oracle(oracle):
    from oracle import interp
    c = cases.lattice(14, 2.0)
    I = interp.Interpreter(casegen.instantiate("lattice_3d", c, (c["N"],)), 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    mom0 = (c["m"][:, None].astype(np.float64) * c["u"]).sum(0)
    for _ in range(3):
        I.step()
    # the lattice is 14 cells of dr = 1, support 2 h = 4: (ulong)(14 / 4) + 6 = 9 cells per axis
    assert list(I.V["n_cells"]) == [9, 9, 9, 729]
    # dt = min(courant h / cs, courant dt_Ma h / |u|): |u| <= 0.01 cs sqrt(3) keeps the fixed value
    assert float(I.V["dt"]) == float(np.float32(0.25 * 2.0 / 40.0))
    assert np.isfinite(I.V["u"]).all() and np.abs(I.V["dudt"]).max() > 0
    # pair forces are antisymmetric: linear momentum is conserved to fp32 rounding
    mom = (I.V["m"][:, None].astype(np.float64) * I.V["u"]).sum(0)
    scale = (np.abs(c["m"][:, None].astype(np.float64) * c["u"])).sum()
    assert np.abs(mom - mom0)[:3].max() < 1e-5 * scale


F_SLAB = ("r", "u", "rho", "dudt", "drhodt")


def _serial(n, hfac, steps, **kw):
    from oracle import interp
    c = cases.lattice(n, hfac, **kw)
    I = interp.Interpreter(casegen.instantiate("lattice_3d", c, (c["N"],)), 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    for _ in range(steps):
        I.step()
    return {k: I.unsorted(k) for k in F_SLAB}, float(I.V["dt"])


def _two_ranks(n, hfac, steps, size=2, fixes=None, **kw):
    import threading
    fixes = fixes or casegen.multi_device_fixes
    from oracle import interp
    tr = interp.LocalTransport(size)
    out, errs = {}, []

    def work(rank):
        try:
            c = cases.lattice_slab(n, hfac, rank, size, **kw)
            txt = fixes(casegen.instantiate("lattice_mpi_3d", c, (c["N"],)))
            I = interp.Interpreter(txt, 3, rank=rank, size=size, transport=tr)
            for k in casegen.STATE_FIELDS:
                I.V[k][...] = c[k]
            for _ in range(steps):
                I.step()
            res = {k: I.unsorted(k) for k in F_SLAB}
            res.update({k + "_dev": I.V[k].copy() for k in F_SLAB + ("imove",)})
            res.update(own=c["own"], imove=I.unsorted("imove"), dt=float(I.V["dt"]), tools=len(I.tools))
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


def test_z_slabs_on_two_ranks_reproduce_the_serial_run(oracle):
    """BASELINE config 5 on two devices (CPU oracle, two threads): the lattice split into z slabs,
    run through cases_xml/lattice_mpi_3d (lattice pipeline + the reference's cfd/MPI.xml and
    cfd/MPI/planes.xml presets, with casegen.multi_device_fixes), reproduces the one-device run of
    lattice_3d: the same dt on both ranks, fields to fp32 summation order (the remote pair terms
    are added after the local ones)."""
    serial, dt = _serial(12, 2.0, 2)
    ranks = _two_ranks(12, 2.0, 2)
    cover = np.concatenate([ranks[r]["own"] for r in range(2)])
    assert np.array_equal(np.sort(cover), np.arange(12 ** 3))       # the slabs partition the lattice
    for r in range(2):
        own = ranks[r]["own"]
        n = len(own)
        assert n == 12 ** 3 // 2 and ranks[r]["tools"] == 79           # 76 + the global-dt all-reduce + the outgoing-mask backup / restore
        assert ranks[r]["dt"] == dt
        assert (ranks[r]["imove"][:n] == 1).all() and (ranks[r]["imove"][n:] == -255).all()
        for k, tol in (("r", 5e-7), ("u", 2e-6), ("rho", 5e-7), ("dudt", 1e-4), ("drhodt", 5e-6)):
            a = serial[k][own].astype(np.float64)
            b = ranks[r][k][:n].astype(np.float64)
            assert np.abs(a - b).max() <= tol * np.abs(a).max(), (r, k, np.abs(a - b).max() / np.abs(a).max())


def test_adams_bashforth_variant(oracle, tmp_path):
    """lattice_ab_3d: the lattice pipeline with basic/time_scheme/adams_bashforth.xml (38 tools).  Host
    front-end and interpreter agree on it; its first step (order 1) equals the improved-Euler pipeline's,
    later steps differ by the scheme, not by orders of magnitude."""
    from oracle import interp
    c = cases.lattice(12, 2.0)
    txt = casegen.instantiate("lattice_ab_3d", c, (c["N"],))
    p = tmp_path / "ab.xml"
    p.write_text(txt)
    tools = host.Simulation(str(p), dims=3, parse_only=True).tools()
    I = interp.Interpreter(txt, 3)
    assert tools == [(t["name"], t["type"]) for t in I.tools] and len(tools) == 38
    names = [n for n, _ in tools]
    assert names.index("sort adams-bashforth") < names.index("Sort") < names.index("corrector") < \
        names.index("backup adams-bashforth") < names.index("Corrector")
    assert I.defs.get("TSCHEME_ADAMS_BASHFORTH_STEPS") == "5u"
    J = interp.Interpreter(casegen.instantiate("lattice_3d", c, (c["N"],)), 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
        J.V[k][...] = c[k]
    for step in range(6):
        I.step()
        J.step()
        diff = np.abs(I.unsorted("u") - J.unsorted("u")).max()
        assert (diff == 0) if step == 0 else (0 < diff < 0.05 * np.abs(J.unsorted("u")).max()), (step, diff)


def test_z_slabs_on_four_ranks_reproduce_the_serial_run(oracle):
    """Interior ranks have a neighbour on both faces (the low_ and high_ instances of
    cfd/MPI/planes.xml both active), as on the 8-GPU runs of config 5: four ranks of 4, 6, 6 and 4 layers
    (planes every (n + 4) / 4 = 6 from z = -2) against the one-device run.  hfac = 1, i.e. a halo of
    2h = 2 layers: the reference's masks hold ONE destination per particle (cfd/MPI/planes.cl:41-99), so
    a slab must be thicker than the two halos it feeds -- with hfac = 2 the middle layers of a 6-layer
    slab would be wanted by both neighbours and reach only one of them."""
    n = 20
    serial, dt = _serial(n, 1.0, 2)
    ranks = _two_ranks(n, 1.0, 2, size=4)
    assert np.array_equal(np.sort(np.concatenate([ranks[r]["own"] for r in range(4)])), np.arange(n ** 3))
    for r in range(4):
        own = ranks[r]["own"]
        m = len(own)
        assert m in (4 * n * n, 6 * n * n) and ranks[r]["dt"] == dt   # planes at z = 4, 10, 16
        assert (ranks[r]["imove"][:m] == 1).all()
        for k, tol in (("r", 5e-7), ("u", 2e-6), ("rho", 5e-7), ("dudt", 1e-4), ("drhodt", 5e-6)):
            a = serial[k][own].astype(np.float64)
            b = ranks[r][k][:m].astype(np.float64)
            assert np.abs(a - b).max() <= tol * np.abs(a).max(), (r, k, np.abs(a - b).max() / np.abs(a).max())


def _gloo_rank(rank, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=2)
    from aquagpusph_b200 import casegen as cg, cases as cs
    from oracle import interp as ip, oracle as O
    O.build()
    c = cs.lattice_slab(12, 2.0, rank, 2)
    txt = cg.multi_device_fixes(cg.instantiate("lattice_mpi_3d", c, (c["N"],)))
    I = ip.Interpreter(txt, 3, rank=rank, size=2, transport=ip.TorchTransport())
    for k in cg.STATE_FIELDS:
        I.V[k][...] = c[k]
    for _ in range(2):
        I.step()
    q.put((rank, {k: np.asarray(I.unsorted(k)) for k in F_SLAB}))
    dist.barrier()
    dist.destroy_process_group()


def test_z_slabs_two_processes_gloo(oracle):
    """world_size = 2 over torch.distributed / gloo (one interpreter rank per process, the arrangement
    of the GPU runs): bit-identical to the two threaded ranks."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_rank, args=(r, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = _two_ranks(12, 2.0, 2)
    for r in range(2):
        for k in F_SLAB:
            assert np.array_equal(got[r][k], want[r][k]), (r, k)


def test_migrating_particles_are_neither_lost_nor_duplicated(oracle):
    """Particles that cross a cut travel through `mpi local sync` + cfd/MPI.cl append / remove.  The
    reference's pipeline hands `remove` the mask mpi-sync has already rewritten, so it parks the
    wrong rows (casegen.multi_device_fixes, (iii)): with the outgoing mask restored, three ranks
    hold exactly the serial run's particles after the crossings, field by field; without it they
    do not.  Fast particles on off-lattice positions: about a tenth of a layer crosses each cut."""
    kw = dict(uscale=0.2, jitter=0.45)
    n, steps, size = 22, 4, 3      # (a slab must be thicker than the two halos it feeds: 26 / 3 > 2 x 4)
    serial, dt = _serial(n, 2.0, steps, **kw)
    ranks = _two_ranks(n, 2.0, steps, size=size, **kw)

    def live(res):
        rows = np.flatnonzero(res["imove_dev"] == 1)
        return rows, res["r_dev"][rows]

    from scipy.spatial import cKDTree
    tree = cKDTree(serial["r"][:, :3].astype(np.float64))
    seen, moved = [], 0
    for r in range(size):
        # one dt for all ranks (all-reduced); against the serial run the velocities it derives from
        # carry the rounding of a different summation order by now
        assert ranks[r]["dt"] == ranks[0]["dt"] and abs(ranks[r]["dt"] - dt) <= 1e-6 * dt
        rows, pos = live(ranks[r])
        d, j = tree.query(pos[:, :3].astype(np.float64))
        assert d.max() < 1e-5, (r, d.max())
        seen.append(j)
        moved += len(set(j.tolist()) ^ set(ranks[r]["own"].tolist()))
        for k, tol in (("u", 2e-6), ("rho", 2e-6), ("dudt", 2e-5), ("drhodt", 1e-5)):
            a = serial[k][j].astype(np.float64)
            b = ranks[r][k + "_dev"][rows].astype(np.float64)
            assert np.abs(a - b).max() <= tol * np.abs(serial[k]).max(), (r, k)
    seen = np.concatenate(seen)
    assert moved > 20, "no particle crossed a cut: the test does not test"
    assert len(seen) == n ** 3 and len(np.unique(seen)) == n ** 3, "particles lost or duplicated"

    # the reference's order of things does lose / duplicate them
    def without_iii(txt):
        txt = casegen.multi_device_fixes(txt)
        return re.sub(r'\s*<Tool [^>]*name="mpi sent mask (backup|restore)"[^>]*/>', "", txt)
    bad = _two_ranks(n, 2.0, steps, size=size, fixes=without_iii, **kw)
    n_bad = sum(int((bad[r]["imove_dev"] == 1).sum()) for r in range(size))
    rows = [live(bad[r])[1] for r in range(size)]
    d, j = tree.query(np.concatenate(rows)[:, :3].astype(np.float64))
    assert n_bad != n ** 3 or len(np.unique(j[d < 1e-3])) != n ** 3
