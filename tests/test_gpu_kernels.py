"""GPU parity: every hot-path script kernel through the Kernel-tool C-ABI
(arguments bound by name) vs the oracle on the same sorted inputs.

Tolerance (fp32): |gpu - oracle| <= 2e-6 * max|oracle| + 2e-5 * |oracle| per
array; element-wise kernels built with -fmad=false and index outputs bit-exact."""
import numpy as np
import pytest

import cases
import pipeline
from aquagpusph_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", [3, 2])
@pytest.mark.parametrize("dims,n,hfac", [(3, 14, 2.0), (2, 60, 3.0), (3, 10, 3.0), (2, 40, 4.0)])
def test_sweeps_match_oracle(oracle, dims, n, hfac, engine):
    """engine 3 = CTA-shared tiles with deferred bodies (the default above ~1 M particles in
    2-D, always in 3-D), engine 2 = per-warp tiles in the reference's visiting order."""
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    ref = pipeline.oracle_sweeps(s)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    try:
        assert _lib.lib().aqc_sweep_engine_select(engine) == engine
        got = pipeline.cuda_sweeps(ctx, s)
    finally:
        _lib.lib().aqc_sweep_engine_select(-1)
    ctx.close()
    rep = pipeline.compare(ref, got)
    bad = [r for r in rep if not r[3]]
    assert not bad, "\n".join("%s: max err %.3e (scale %.3e)" % r[:3] for r in bad)
    # the case must actually exercise the kernels
    assert np.abs(ref["grad_p"]).max() > 0 and np.abs(ref["grad_w_bi"]).max() > 0
    assert np.abs(ref["lap_p"]).max() > 0 and (ref["n_neighs"] > 0).any()


@pytest.mark.parametrize("dims,n,hfac,scale", [(3, 10, 3.0, 3.5), (3, 12, 2.0, 0.45), (2, 50, 3.0, 4.0),
                                               (2, 60, 4.0, 0.3)])
def test_sweeps_sparse_and_dense_cells(oracle, dims, n, hfac, scale):
    """The same kernels on stretched positions (0-2 particles per cell: every warp spans many
    cells and several group passes) and on compressed ones (thousands of candidates per cell:
    long runs, many ring rounds): the CTA-shared walk must still visit exactly the
    reference's pairs."""
    case = cases.dam_break(dims, n, hfac)
    for k in ("r",):
        case[k] = (case[k] * np.float32(scale)).astype(np.float32)
    case["domain_min"] = (np.asarray(case["domain_min"]) * np.float32(scale) - 1).astype(np.float32)
    case["domain_max"] = (np.asarray(case["domain_max"]) * np.float32(scale) + 1).astype(np.float32)
    s = pipeline.oracle_linklist_and_sort(case)
    ref = pipeline.oracle_sweeps(s)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    try:
        _lib.lib().aqc_sweep_engine_select(3)
        got = pipeline.cuda_sweeps(ctx, s)
    finally:
        _lib.lib().aqc_sweep_engine_select(-1)
    ctx.close()
    # sums of thousands of terms in another order: twice the usual absolute band
    bad = [r for r in pipeline.compare(ref, got) if not r[3] and r[1] > 5e-6 * r[2]]
    assert not bad, "\n".join("%s: max err %.3e (scale %.3e)" % r[:3] for r in bad)
    # the pair SET itself, exactly: fluid neighbours within the support counted by the
    # default engine and by the order-preserving per-warp engine
    L = _lib.lib()
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    st = pipeline.CudaState(ctx, s)
    counts = []
    try:
        for eng in (3, 2):
            assert L.aqc_sweep_engine_select(eng) == eng
            st.v["n_pairs"] = ctx.zeros(st.v["N"], np.uint32)
            st.run("aqua/diag.cl", "count_pairs")
            counts.append(st.v["n_pairs"].get())
    finally:
        L.aqc_sweep_engine_select(-1)
    ctx.close()
    assert counts[0].sum() > 0 and np.array_equal(counts[0], counts[1])


@pytest.mark.parametrize("dims,n,hfac", [(3, 14, 2.0), (2, 60, 3.0)])
def test_fused_fluid_sweep_equals_its_members(oracle, dims, n, hfac):
    """aqc_launch_fused: Shepard + Interactions + deltaSPH full + lapp in one pass must give
    what the four sweeps give one after the other: same pairs, same order, same expressions;
    only the compiler's FMA contraction may differ between the two kernels (a few ulp of the
    largest term), rows no member writes stay untouched."""
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    sets = [
        [("cfd/Shepard.cl", "entry"), ("cfd/Interactions.cl", "entry"), ("cfd/deltaSPH.cl", "full"),
         ("cfd/deltaSPH.cl", "lapp")],
        [("cfd/Shepard.cl", "entry"), ("cfd/Interactions.cl", "entry")],
        [("cfd/Interactions.cl", "entry"), ("cfd/deltaSPH.cl", "full"), ("cfd/deltaSPH.cl", "lapp")],
    ]
    outs = ("shepard", "grad_p", "lap_u", "div_u", "lap_p_corr", "lap_p")
    for members in sets:
        a = pipeline.CudaState(ctx, s)
        b = pipeline.CudaState(ctx, s)
        for st in (a, b):
            st.run("basic/EOS.cl")
            for k in outs:          # sentinel: rows the kernels do not write must stay untouched
                st.ctx.fill(st.v[k], np.full(st.v[k].elem_bytes // 4, 7.5, np.float32).tobytes())
        for sc, en in members:
            a.run(sc, en)
        ctx.launch_fused(members, b.v)
        for k in outs:
            x, y = a.get(k).astype(np.float64), b.get(k).astype(np.float64)
            assert np.array_equal(x == 7.5, y == 7.5), (members, k)
            assert np.abs(x - y).max() <= 2e-6 * np.abs(x).max(), (members, k)
    L = _lib.lib()
    import ctypes as C
    ids = (C.c_int * 2)(ctx.lookup("cfd/Rates.cl", "entry"), ctx.lookup("cfd/Interactions.cl", "entry"))
    assert L.aqc_fused_lookup(ids, 2, dims) < 0
    ctx.close()


@pytest.mark.parametrize("dims,n,hfac,scale", [(3, 14, 2.0, 1.0), (3, 10, 3.0, 1.0), (2, 60, 3.0, 1.0),
                                               (3, 10, 3.0, 3.5), (3, 12, 2.0, 0.45)])
def test_pair_mask_cache_is_bit_identical(oracle, dims, n, hfac, scale):
    """aqc_pairs_cache_enable: MLS, Shepard, Interactions, delta-SPH full / lapp / lapp_corr read the
    hit masks one builder pass stored instead of filtering.  Every selected pair is still tested
    exactly and the hits are consumed in the same order, so every output must be BIT-identical
    with and without the cache -- also on stretched positions (a CTA spans dozens of cells: more
    passes than the pass table holds, the sweeps fall back to filtering) and compressed ones."""
    case = cases.dam_break(dims, n, hfac)
    if scale != 1.0:
        case["r"] = (case["r"] * np.float32(scale)).astype(np.float32)
        case["domain_min"] = (np.asarray(case["domain_min"]) * np.float32(scale) - 1).astype(np.float32)
        case["domain_max"] = (np.asarray(case["domain_max"]) * np.float32(scale) + 1).astype(np.float32)
    s = pipeline.oracle_linklist_and_sort(case)
    L = _lib.lib()
    out, stats = [], None
    try:
        assert L.aqc_sweep_engine_select(3) == 3
        for cache in (False, True):
            ctx = _lib.Context(0, dims=dims, h=case["h"])
            ctx.pairs_cache(cache)
            out.append(pipeline.cuda_sweeps(ctx, s))
            if cache:
                stats = ctx.pairs_cache_stats()
            ctx.close()
    finally:
        L.aqc_sweep_engine_select(-1)
    for k in out[0]:
        a, b = np.asarray(out[0][k]), np.asarray(out[1][k])
        assert a.tobytes() == b.tobytes(), k
    # one build serves the sweeps up to the first kernel that writes r (PST, at the end)
    assert stats["builds"] >= 1 and (scale == 3.5 or stats["hits"] >= 5), stats


def test_pair_mask_cache_is_dropped_when_its_inputs_change(oracle):
    """Writing r, imove or the link-list through the library must invalidate the masks: the
    sweep after the write sees the new positions (bit-identical to a context without cache)."""
    dims = 3
    case = cases.dam_break(dims, 12, 2.0)
    s = pipeline.oracle_linklist_and_sort(case)
    rng = np.random.default_rng(5)
    r2 = s["r"].copy()
    r2[:, :dims] += (rng.random((s["N"], dims), dtype=np.float32) - 0.5) * np.float32(0.8 * case["h"])
    mv2 = s["imove"].copy()
    fl = np.flatnonzero(mv2 == 1)
    mv2[fl[::7]] = -1  # some fluid particles leave the i and j sets
    L = _lib.lib()
    res = []
    try:
        assert L.aqc_sweep_engine_select(3) == 3
        for cache in (False, True):
            ctx = _lib.Context(0, dims=dims, h=case["h"])
            ctx.pairs_cache(cache)
            st = pipeline.CudaState(ctx, s)
            st.run("basic/EOS.cl")
            got = []
            for step in range(4):
                if step == 1:
                    st.set("r", r2)                     # aqc_memcpy_h2d
                elif step == 2:
                    st.set("imove", mv2)
                elif step == 3:
                    st.run("cfd/Boundary/BIe/PST.cl")  # a kernel with r among its outputs
                st.run("cfd/Interactions.cl")
                st.run("cfd/Shepard.cl")
                got += [st.get("grad_p"), st.get("div_u"), st.get("shepard"), st.get("r")]
            res.append(got)
            if cache:
                stats = ctx.pairs_cache_stats()
            ctx.close()
    finally:
        L.aqc_sweep_engine_select(-1)
    for a, b in zip(*res):
        assert a.tobytes() == b.tobytes()
    assert not np.array_equal(res[0][0], res[0][4]) and not np.array_equal(res[0][4], res[0][8])
    assert stats["builds"] == 5 and stats["hits"] == 8, stats  # step 0 builds twice (Shepard adds i classes)


@pytest.mark.parametrize("engine", [3, 2])
@pytest.mark.parametrize("dims,n,hfac", [(3, 10, 3.0), (2, 40, 4.0)])
def test_sweeps_match_the_reference_outputs(dims, n, hfac, engine):
    """The same sequence against tests/golden/hotpath_sweeps_outputs.npz: what the reference's OWN scripts produced
    for these states (tests/golden/make_golden_hotpath.py) -- no oracle in between.  (The oracle equals those
    outputs bit for bit, tests/test_oracle_golden.py, so the bar is the one of test_sweeps_match_oracle.)"""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_sweeps_outputs.npz"))
    pre = "sweeps_%dD_" % dims
    ref = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    if "dt" in ref:
        ref["dt"] = np.float32(ref["dt"])
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)      # (the cell-sorted INPUT state: link-list of the test infrastructure)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    try:
        assert _lib.lib().aqc_sweep_engine_select(engine) == engine
        got = pipeline.cuda_sweeps(ctx, s)
    finally:
        _lib.lib().aqc_sweep_engine_select(-1)
    ctx.close()
    assert len(ref) >= 30 and set(ref) == set(got)
    bad = [r for r in pipeline.compare(ref, got) if not r[3]]
    assert not bad, "\n".join("%s: max err %.3e (scale %.3e)" % r[:3] for r in bad)
