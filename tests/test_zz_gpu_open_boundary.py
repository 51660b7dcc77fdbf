"""GPU parity of the open-boundary kernels written by hand after round 2's GPU budget was spent
(cfd/Boundary/Inlet/Inlet.cl::feed / rates, Outlet/Outlet.cl::rates / feed, Portal/Mirror.cl::mirror / unmirror /
teleport; presets cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml) through the Kernel-tool C-ABI,
against the oracle AND against the committed outputs of the reference's own scripts
(tests/golden/open_boundary_outputs.npz, made by make_golden_open_boundary.py where the reference tree is).

In a file of its own that sorts after every other GPU suite: its first run on a B200 is the driver's, and nothing
may hide behind it under -x.  The CPU halves are tests/test_oracle_vs_reference.py (restatement == scripts),
tests/test_oracle_golden.py (restatement == committed outputs) and tests/test_presets_host_emulation.py (the
CUDA kernel bodies and their launchers' arithmetic, compiled for the host == restatement)."""
import os

import numpy as np
import pytest

import open_boundary_common as ob
from aquagpusph_b200 import _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "open_boundary_outputs.npz")


@pytest.mark.parametrize("dims", [2, 3])
def test_open_boundary_kernels(oracle, dims):
    """Element-wise, no contraction (-fmad=false), IEEE + - * / only: bit-exact after every kernel."""
    import oracle.oracle as O
    case, v = ob.state(dims)
    D = O.make_defs(dims, case["h"])
    G = np.load(GOLD)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    o = ob.args_of(v)
    d = {k: (ctx.array(x) if isinstance(x, np.ndarray) and k not in ob.SCALARS else x) for k, x in v.items()}
    d["refd"] = ctx.array(v["refd"])
    for step, key in enumerate(ob.STEPS):
        # (the Kernel tool's global size is N for every one of them; the inlet launcher trims it itself)
        ctx.launch(key[0], key[1], d)
        ob.oracle_step(oracle, D, dims, key, o)
        for k in ob.WRITES[key]:
            got = d[k].get()
            assert got.tobytes() == o[k].tobytes(), (key, k)
            assert got.tobytes() == G["%dD_step%d_%s" % (dims, step, k)].tobytes(), (key, k)
    final = {k: d[k].get() for k in o if isinstance(o[k], np.ndarray) and k not in ob.SCALARS}
    final["n_cells"], final["inlet_N"] = v["n_cells"], v["inlet_N"]
    final["N"], final["nbuffer"] = v["N"], v["nbuffer"]
    ob.checks(v, final, dims)
    # a starving flag of 0: no launch, nothing touched
    before = ctx.launch_count()
    d2 = dict(d)
    d2["inlet_starving"] = 0
    ctx.launch(ob.INLET, "feed", d2)
    assert ctx.launch_count() == before
    ctx.close()


@pytest.mark.parametrize("engine", [3, 2])
@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0), (2, 40, 4.0)])
def test_interactions_morris_laplacian_sweep(oracle, dims, n, hfac, engine):
    """cfd/Interactions.cl::entry under <Define name="__LAP_FORMULATION__" value="__LAP_MORRIS__"/> (examples/2D/
    taylor_green, cylinder_inside_channel; Interactions.cl:130-131) through the Kernel-tool C-ABI on both sweep
    engines vs the oracle, which is bit-identical to the reference's script compiled with that definition
    (tests/test_oracle_vs_reference.py).  Tolerance of the sweep tests:
    |gpu - oracle| <= 2e-6 max|oracle| + 2e-5 |oracle|; the rows of the other particle classes untouched.  The
    CPU half is tests/test_sweep_policy_host_emulation.py::test_interactions_policies_match_the_oracle."""
    import cases
    import pipeline
    from test_oracle_vs_reference import morris_inputs
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = morris_inputs(s)
    want = {k: x[k].copy() for k in ("grad_p", "lap_u", "div_u")}
    oracle.call("interactions_morris", oracle.make_defs(dims, s["h"]), pipeline._ll(s), s["imove"], s["r"], x["u"],
                s["rho"], s["m"], x["p"], want["grad_p"], want["lap_u"], want["div_u"])
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    assert _lib.lib().aqc_set_define(ctx.h, b"__LAP_FORMULATION__", b"__LAP_MORRIS__") == 0
    c = pipeline.CudaState(ctx, s)
    for k in ("u", "p", "grad_p", "lap_u", "div_u"):
        c.set(k, x[k])
    try:
        assert _lib.lib().aqc_sweep_engine_select(engine) == engine
        c.run("cfd/Interactions.cl")
        got = {k: c.get(k) for k in want}
        # the fused groups hold the Monaghan build: refused under this definition, loudly
        with pytest.raises(_lib.AquaError, match="__LAP_MONAGHAN__"):
            ctx.launch_fused([("cfd/Shepard.cl", "entry"), ("cfd/Interactions.cl", "entry")], c.v)
    finally:
        _lib.lib().aqc_sweep_engine_select(-1)
    ctx.close()
    fl = s["imove"] == 1
    for k in want:
        a, b = want[k].astype(np.float64), got[k].astype(np.float64)
        assert np.isfinite(b).all(), k
        assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a[fl]).max() + 2e-5 * np.abs(a)), (k, np.abs(a - b).max())
        assert np.array_equal(got[k][~fl], x[k][~fl]) and np.abs(want[k][fl] - x[k][fl]).max() > 0, k


@pytest.mark.parametrize("morris", [0, 1])
@pytest.mark.parametrize("dims,n,hfac", [(2, 60, 3.0), (3, 24, 1.3), (2, 80, 4.0)])
def test_portal_sweeps(oracle, dims, n, hfac, morris):
    """cfd/Boundary/Portal/Shepard.cl::entry and Portal/Interactions.cl::entry (under both Laplacian definitions;
    preset cfd/portal.xml, examples/2D/taylor_green) through the Kernel-tool C-ABI on a state
    Portal/Mirror.cl::mirror prepared (mirrored particles keep their rows, icell names their new cells) vs the
    oracle, which is bit-identical to the reference's scripts.  The sweeps ADD to what the arrays hold: the sweep
    tests' tolerance on the added part, |d_gpu - d_oracle| <= 2e-6 max(|d_oracle|, |array|) + 2e-5 |d_oracle|;
    rows that are not mirrored keep their bits.  CPU half: tests/test_sweep_policy_host_emulation.py::
    test_portal_policies_match_the_oracle."""
    import pipeline
    case, s, D, x = ob.portal_sweep_state(oracle, dims, n, hfac)
    want = ob.portal_sweeps_oracle(oracle, s, D, x, morris)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    if morris:
        assert _lib.lib().aqc_set_define(ctx.h, b"__LAP_FORMULATION__", b"__LAP_MORRIS__") == 0
    c = pipeline.CudaState(ctx, s)
    for k in ("r", "icell", "u", "p", "grad_p", "lap_u", "div_u", "shepard"):
        c.set(k, x[k])
    c.v["imirrored"] = ctx.array(x["imirrored"])
    c.run("cfd/Boundary/Portal/Shepard.cl")
    c.run("cfd/Boundary/Portal/Interactions.cl")
    got = {k: c.get(k) for k in want}
    ctx.close()
    for k in want:
        rows = ob.portal_rows(s, x, k)
        add_w = want[k].astype(np.float64) - x[k]
        add_g = got[k].astype(np.float64) - x[k]
        assert np.isfinite(got[k]).all() and np.array_equal(got[k][~rows], x[k][~rows]), k
        tol = 2e-6 * max(np.abs(add_w).max(), np.abs(x[k]).max()) + 2e-5 * np.abs(add_w)
        assert np.all(np.abs(add_w - add_g) <= tol), (k, np.abs(add_w - add_g).max())
