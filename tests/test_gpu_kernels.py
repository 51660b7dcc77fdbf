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


@pytest.mark.parametrize("dims,n,hfac", [(3, 14, 2.0), (2, 60, 3.0), (3, 10, 3.0), (2, 40, 4.0)])
def test_sweeps_match_oracle(oracle, dims, n, hfac):
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    ref = pipeline.oracle_sweeps(s)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    got = pipeline.cuda_sweeps(ctx, s)
    ctx.close()
    rep = pipeline.compare(ref, got)
    bad = [r for r in rep if not r[3]]
    assert not bad, "\n".join("%s: max err %.3e (scale %.3e)" % r[:3] for r in bad)
    # the case must actually exercise the kernels
    assert np.abs(ref["grad_p"]).max() > 0 and np.abs(ref["grad_w_bi"]).max() > 0
    assert np.abs(ref["lap_p"]).max() > 0 and (ref["n_neighs"] > 0).any()


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
