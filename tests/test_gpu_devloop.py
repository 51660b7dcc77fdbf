"""Device-side loops: a `while` of the pipeline as one CUDA graph with a WHILE
conditional node (csrc/devloop.cu, host/devloop.cpp) against the SAME loop run tool by tool on the
host, which is what the reference does (Conditional.cpp:85-96, SetScalar.cpp:146-195,
Reduction.cpp:205-258).

Bar: bit-exact -- the recorded body launches the same kernels in the same order, the scalar programs
compute in IEEE double one operation at a time like the host's evaluator (pinned on the CPU by
tests/test_host_cpu.py::test_svm_programs_equal_the_host_evaluator), so every array, every scalar and
every report line must be identical, not close."""
import os
import struct

import numpy as np
import pytest

from aquagpusph_b200 import _lib, cases, casegen, host
from aquagpusph_b200._lib import (AQS_ADD, AQS_FOLD, AQS_IMM, AQS_LOAD, AQS_LT, AQS_MUL, AQS_SETCOND,
                                  AQS_SNAP, AQS_STORE, AQS_ASSERT, AQS_RECOND)

pytestmark = pytest.mark.gpu


class _Ptr:
    def __init__(self, p):
        self.ptr = p


def _table(**slots):
    """16-byte slots in keyword order: name=(fmt, values...)"""
    b = b""
    for fmt, *vals in slots.values():
        raw = struct.pack("<" + fmt, *vals)
        b += raw + b"\0" * (16 - len(raw))
    return b


def test_loop_c_abi_counts_reduces_and_reads_scalars_on_the_device():
    """aqc_loop_*: fill + reduction + a kernel reading a loop scalar + scalar programs, five passes."""
    ctx = _lib.Context(0, dims=3, h=0.1)
    n = 70001
    rng = np.random.default_rng(3)
    a_h = rng.random(n, dtype=np.float32)
    a = ctx.array(a_h)
    b = ctx.zeros(n, np.float32)
    imove = ctx.array(np.ones(n, np.int32))
    dudt_in_h = rng.random((n, 4), dtype=np.float32)
    dudt_h = rng.random((n, 4), dtype=np.float32)
    drho_in_h = rng.random(n, dtype=np.float32)
    drho_h = rng.random(n, dtype=np.float32)
    dudt_in, dudt = ctx.array(dudt_in_h), ctx.array(dudt_h)
    drho_in, drho = ctx.array(drho_in_h), ctx.array(drho_h)
    # table: k @0 (u), x @16 (f), S @32 (f), raw @48, null @64, relax @80 (f)
    tab = _table(k=("I", 0), x=("f", 1.0), S=("f", 0.0), raw=("f", 0.0), null=("f", 0.5), relax=("f", 0.8))
    loop = ctx.loop(len(tab), hist_rows=8)
    cond = [(AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 5), (AQS_LT,), (AQS_SETCOND,)]
    before = ctx.launch_count()
    loop.begin(cond)
    ctx.fill(b, np.float32(2.0).tobytes())
    ctx.reduce(_lib.OP_SUM, a, out_dev=_Ptr(loop.table() + 48), host=False)
    ctx.launch("basic/time_scheme/midpoint.cl", "relax",
               dict(imove=imove, dudt_in=dudt_in, dudt=dudt, drhodt_in=drho_in, drhodt=drho, N=n,
                    relax_midpoint=-1.0),                     # the by-value fallback must NOT be used
               dev_scalars={"relax_midpoint": loop.table() + 80})
    loop.svm([(AQS_FOLD, 32, 48, 64, _lib.OP_SUM + 4 * 0 + 16 * 1),
              (AQS_LOAD, 16, "f"), (AQS_IMM, 0, 0, 0, 2.0), (AQS_MUL,), (AQS_STORE, 16, "f"),
              (AQS_LOAD, 80, "f"), (AQS_IMM, 0, 0, 0, 0.5), (AQS_MUL,), (AQS_STORE, 80, "f"),
              (AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 1), (AQS_ADD,), (AQS_STORE, 0, "u"),
              (AQS_SNAP, 7)] + cond)
    loop.end()
    assert ctx.launch_count() == before          # recording launched nothing
    nodes, rec_ms, inst_ms = loop.stats()
    assert nodes >= 5
    hdr, out, rows = loop.run(tab)
    assert (hdr.iters, hdr.snaps, hdr.error, hdr.cond) == (5, 5, 0, 0)
    k, = struct.unpack_from("<I", out, 0)
    x, S, raw, null, relax = (struct.unpack_from("<f", out, o)[0] for o in (16, 32, 48, 64, 80))
    assert k == 5 and x == 32.0 and null == 0.5 and relax == np.float32(0.8) * np.float32(0.5) ** 5
    # the reduction's own fixed-order sum, folded with the null value in fp32
    want = ctx.reduce(_lib.OP_SUM, a)
    assert raw == want and S == np.float32(want + np.float32(0.5))
    assert [r[0] for r in rows] == [7] * 5
    assert [struct.unpack_from("<I", r[1], 0)[0] for r in rows] == [1, 2, 3, 4, 5]
    assert np.all(b.get() == 2.0)
    # the kernel read relax from the table: 0.8, 0.4, 0.2, 0.1, 0.05 -- the same five launches by value
    d2, r2 = ctx.array(dudt_h), ctx.array(drho_h)
    f = np.float32(0.8)
    for _ in range(5):
        ctx.launch("basic/time_scheme/midpoint.cl", "relax",
                   dict(imove=imove, dudt_in=dudt_in, dudt=d2, drhodt_in=drho_in, drhodt=r2, N=n,
                        relax_midpoint=float(f)))
        f = np.float32(f * np.float32(0.5))
    assert np.array_equal(dudt.get(), d2.get()) and np.array_equal(drho.get(), r2.get())
    # 1 entry program + 5 x (fill, 2 reduction kernels, relax, program) -- plus what the check above launched
    assert ctx.launch_count() - before == 1 + 5 * 5 + 2 + 5
    # a second run of the same recording, other initial values, iteration cap
    hdr, out, rows = loop.run(_table(k=("I", 3), x=("f", 1.0), S=("f", 0.0), raw=("f", 0.0),
                                     null=("f", 0.0), relax=("f", 1.0)))
    assert hdr.iters == 2 and struct.unpack_from("<f", out, 16)[0] == 4.0
    hdr, out, rows = loop.run(tab, max_iters=3)
    assert hdr.iters == 3 and hdr.error == 0x30000
    loop.close()
    # the way the host drives it: the first pass runs directly on the uploaded table (nothing read
    # back), the graph starts from the condition that pass left
    loop = ctx.loop(len(tab), hist_rows=8)
    body = [(AQS_FOLD, 32, 48, 64, _lib.OP_SUM + 4 * 0 + 16 * 1),
            (AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 1), (AQS_ADD,), (AQS_STORE, 0, "u"), (AQS_SNAP, 9)] + cond
    before = ctx.launch_count()
    loop.start(tab)
    for recording in (False, True):
        if recording:
            loop.begin([(AQS_RECOND,)])
        ctx.fill(b, np.float32(3.0).tobytes())
        ctx.reduce(_lib.OP_SUM, a, out_dev=_Ptr(loop.table() + 48), host=False)
        loop.svm(body)
        if recording:
            loop.end()
    assert ctx.launch_count() - before == 4      # the direct pass only
    hdr, out, rows = loop.run()
    assert (hdr.iters, hdr.snaps, hdr.error, hdr.cond) == (4, 5, 0, 0)      # 1 direct + 4 graph passes
    assert struct.unpack_from("<I", out, 0)[0] == 5 and struct.unpack_from("<f", out, 32)[0] == S
    assert [struct.unpack_from("<I", r[1], 0)[0] for r in rows] == [1, 2, 3, 4, 5]
    assert ctx.launch_count() - before == 4 + 1 + 4 * 4
    # a recording that failed: run() hands back what the direct pass left
    loop.start(tab)
    loop.svm(body)
    loop.begin([(AQS_RECOND,)])
    with pytest.raises(_lib.AquaError, match="synchronises"):
        ctx.reduce(_lib.OP_SUM, a)
    loop.abort()
    hdr, out, rows = loop.run()
    assert (hdr.iters, hdr.snaps, hdr.cond) == (1, 1, 1) and struct.unpack_from("<I", out, 0)[0] == 1
    loop.close()
    ctx.close()


def test_recording_refuses_what_synchronises_and_leaves_the_context_usable():
    ctx = _lib.Context(0, dims=3, h=0.1)
    a = ctx.array(np.arange(1000, dtype=np.float32))
    tab = _table(k=("I", 0))
    loop = ctx.loop(len(tab))
    cond = [(AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 2), (AQS_LT,), (AQS_SETCOND,)]
    loop.begin(cond)
    with pytest.raises(_lib.AquaError, match="synchronises"):
        ctx.reduce(_lib.OP_SUM, a)          # reads the result back
    loop.abort()
    assert ctx.reduce(_lib.OP_SUM, a) == np.float32(499500.0)
    # a body that never sets the condition cannot end: refused
    loop.begin(cond)
    ctx.fill(a, np.float32(1.0).tobytes())
    with pytest.raises(_lib.AquaError, match="never sets the loop condition"):
        loop.end()
    assert ctx.reduce(_lib.OP_SUM, a) == np.float32(499500.0)   # the fill was recorded, never run
    # an assertion that fails on the device
    loop.begin(cond)
    loop.svm([(AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 1), (AQS_LT,), (AQS_ASSERT, 42),
              (AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 1), (AQS_ADD,), (AQS_STORE, 0, "u")] + cond)
    loop.end()
    hdr, out, rows = loop.run(tab)
    assert hdr.error == 0x10000 + 42 and hdr.iters == 2
    # with a kernel that cannot take the scalar from the device
    with pytest.raises(_lib.AquaError, match="cannot be read from device memory"):
        ctx.launch("basic/EOS.cl", "entry",
                   dict(iset=a, imove=a, rho=a, p=a, refd=a, N=10, cs=1.0, p0=0.0),
                   dev_scalars={"cs": loop.table()})
    loop.close()
    ctx.close()


def _run(template, case, nset, ov, steps, device_loops, monkeypatch, transform=None, dims=3):
    monkeypatch.setenv("AQUA_DEVICE_LOOPS", "1" if device_loops else "0")
    sim = casegen.load(template, case, nset, ov, keep_reports=True, transform=transform)
    out = {"loops": sim.device_loops(), "branch": sim.device_loop_branch_tools()}
    per_step = []
    for _ in range(steps):
        sim.step(1)
        per_step.append((int(sim.scalar("iter_midpoint", np.uint32)), float(sim.scalar("dt")),
                         float(sim.scalar("Residual_midpoint")), float(sim.scalar("relax_midpoint"))))
    out["per_step"] = per_step
    for k in ("r", "u", "rho", "p", "dudt", "drhodt", "dudt_in", "drhodt_in"):
        out[k] = sim.download(k)
    for k, n in (("Force_p", 4), ("Moment_p", 4), ("Force_elastic", 4)):
        try:
            out[k] = sim.scalar(k, np.float32, n if dims == 3 or k == "Moment_p" else 2).copy()
        except host.HostError:
            pass
    out["stats"] = sim.device_loop_stats()
    types = dict(sim.tools())
    out["used"] = {name: used for name, used, _ in sim.tool_times() if types[name] not in ("while", "end")}
    out["why"] = [sim.loop_host_reason(i) for i, (_, ty) in enumerate(sim.tools()) if ty == "while"]
    d = os.path.dirname(sim.xml_path)
    sim.close()
    p = os.path.join(d, "midpoint.out")
    out["report"] = open(p).read() if os.path.exists(p) else None
    return out


def _same(a, b, keys):
    for k in keys:
        if k in a:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("n,maxiter", [(6000, 30), (9000, 4)])
def test_dam_break_midpoint_loop_on_the_device_equals_the_host_loop(n, maxiter, monkeypatch):
    """The 116-tool 3-D dam break (46 tools between `midpoint loop` and its end: fused sweeps,
    reductions, relaxation set_scalars, the residual report, a kernel reading relax_midpoint)."""
    host.set_log_level(3)
    case = cases.spheric2_dam_break(n, 3.0, seed=11)
    nset = (case["N"] - 8, 8)
    ov = {"iter_midpoint_max": maxiter}
    H = _run("spheric2_dambreak_3d", case, nset, ov, 5, False, monkeypatch)
    H2 = _run("spheric2_dambreak_3d", case, nset, ov, 5, False, monkeypatch)
    D = _run("spheric2_dambreak_3d", case, nset, ov, 5, True, monkeypatch)
    assert H["loops"] == 0 and H["why"] == ["AQUA_DEVICE_LOOPS=0"]
    assert D["loops"] == 1 and D["why"] == [""], D["why"]
    # the delta-SPH correction sweep (and what only it depends on) overlaps the boundary chain
    assert D["branch"] >= 2 and H["branch"] == 0
    keys = ("r", "u", "rho", "p", "dudt", "drhodt", "dudt_in", "drhodt_in", "Force_p", "Moment_p",
            "Force_elastic")
    _same(H, H2, keys)                      # the host path repeats itself bit for bit ...
    assert H["per_step"] == H2["per_step"]
    _same(H, D, keys)                       # ... and the device loop gives the same bits
    assert H["per_step"] == D["per_step"]
    assert H["report"] is not None and H["report"] == D["report"]
    # it did run there: once per step, every pass but the first
    assert H["used"] == D["used"], {k: (H["used"][k], D["used"][k]) for k in H["used"] if H["used"][k] != D["used"][k]}
    # it did run there: once per step, every pass of it
    assert D["stats"] == (5, D["used"]["midpoint relax"]) and D["used"]["midpoint relax"] >= 10


def test_tld_midpoint_loop_on_the_device_equals_the_host_loop(monkeypatch):
    """BASELINE config 4's pipeline (2-D, five passes per step, moving tank, force / energy reports)."""
    host.set_log_level(3)
    case = cases.spheric9_tld_2d(3000, 4.0, seed=5)
    nset = (case["n_set0"], case["n_set1"])
    tr = casegen.prescribed_roll(0.05, 0.2)
    ov = {"Residual_midpoint_max": "0.0"}
    H = _run("spheric9_tld_2d", case, nset, ov, 3, False, monkeypatch, transform=tr, dims=2)
    D = _run("spheric9_tld_2d", case, nset, ov, 3, True, monkeypatch, transform=tr, dims=2)
    assert D["loops"] == 1 and D["why"] == [""], D["why"]
    _same(H, D, ("r", "u", "rho", "p", "dudt", "drhodt", "dudt_in", "drhodt_in", "Force_p", "Moment_p",
                 "Force_elastic"))
    assert H["per_step"] == D["per_step"] and [p[0] for p in D["per_step"]] == [5, 5, 5]
    assert H["report"] == D["report"]
    assert D["stats"] == (3, 15)            # five passes in every step
    assert H["used"] == D["used"]


def test_lanes_are_ordered_by_their_events():
    """aqc_lane_select / aqc_lane_event / aqc_lane_wait: work on the branch lane starts behind the event it
    waits for and lane 0 goes on behind the branch's; without a wait the lanes are independent queues."""
    import ctypes as C
    L = _lib.lib()
    ctx = _lib.Context(0, dims=3, h=0.1)
    n = 1 << 25                                     # 128 MB per array: a fill takes tens of microseconds
    a, b, c = ctx.zeros(n, np.float32), ctx.zeros(n, np.float32), ctx.zeros(n, np.float32)
    ev1, ev2 = C.c_void_p(), C.c_void_p()
    for rep in range(3):
        v = np.float32(rep + 1.0)
        ctx.fill(a, v.tobytes())                    # lane 0
        ctx._chk(L.aqc_lane_event(ctx.h, C.byref(ev1)))
        ctx._chk(L.aqc_lane_select(ctx.h, 1))
        ctx._chk(L.aqc_lane_wait(ctx.h, ev1))
        ctx.copy(b, a)                              # lane 1, behind the fill
        ctx._chk(L.aqc_lane_event(ctx.h, C.byref(ev2)))
        ctx._chk(L.aqc_lane_select(ctx.h, 0))
        ctx.fill(c, v.tobytes())                    # lane 0, next to the copy
        ctx._chk(L.aqc_lane_wait(ctx.h, ev2))
        ctx.copy(c, b)                              # lane 0, behind the copy on lane 1
        assert ctx.reduce(_lib.OP_MIN, c) == v and ctx.reduce(_lib.OP_MAX, c) == v
        assert ctx.reduce(_lib.OP_MIN, b) == v
    # a recording given up while the branch lane is inside the capture: both streams come back usable
    loop = ctx.loop(16)
    cond = [(AQS_LOAD, 0, "u"), (AQS_IMM, 0, 0, 0, 2), (AQS_LT,), (AQS_SETCOND,)]
    for rep in range(2):
        loop.begin(cond)
        ctx._chk(L.aqc_lane_event(ctx.h, C.byref(ev1)))
        ctx._chk(L.aqc_lane_select(ctx.h, 1))
        ctx._chk(L.aqc_lane_wait(ctx.h, ev1))
        ctx.fill(b, np.float32(-1.0).tobytes())       # recorded on the branch lane, never run
        with pytest.raises(_lib.AquaError, match="synchronises"):
            ctx.reduce(_lib.OP_MIN, b)
        loop.abort()                                  # (selects lane 0 again)
        assert ctx.reduce(_lib.OP_MIN, b) == 3.0
        ctx._chk(L.aqc_lane_select(ctx.h, 1))
        ctx.fill(a, np.float32(9.0).tobytes())        # the branch lane works outside a capture
        ctx._chk(L.aqc_lane_event(ctx.h, C.byref(ev2)))
        ctx._chk(L.aqc_lane_select(ctx.h, 0))
        ctx._chk(L.aqc_lane_wait(ctx.h, ev2))
        assert ctx.reduce(_lib.OP_MIN, a) == 9.0
    loop.close()
    # loops and their recordings belong on lane 0
    loop = ctx.loop(16)
    ctx._chk(L.aqc_lane_select(ctx.h, 1))
    with pytest.raises(_lib.AquaError, match="branch lane"):
        loop.begin([(AQS_LOAD, 0, "u"), (AQS_SETCOND,)])
    ctx._chk(L.aqc_lane_select(ctx.h, 0))
    loop.close()
    ctx.close()
