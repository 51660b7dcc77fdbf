"""GPU parity of the boundary-integral (BI) kernels of BASELINE config 1 (2-D SPHERIC
test 5 dam break; cfd/Boundary/BI/*.cl, cfd/Boundary/ElasticBounce.cl) against the
reference's OWN scripts compiled behind the shim (oracle/_ref), arguments bound by
name; and of the whole 57-tool pipeline against the oracle interpreter."""
import numpy as np
import pytest

import cases
import pipeline
from aquagpusph_b200 import _lib
from oracle import ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]

ORDER_DEP = {"u_bounce", "dudt_bounce"}


@pytest.mark.parametrize("dims,n,hfac", [(2, 60, 3.0), (2, 40, 4.0), (3, 12, 2.0)])
def test_bi_kernels_match_reference_scripts(oracle, dims, n, hfac):
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    want = pipeline.named_bi_sequence(pipeline.RefState(ref.Ref(dims, case["h"]), s))
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    got = pipeline.named_bi_sequence(pipeline.CudaState(ctx, s))
    ctx.close()
    bad = []
    for k, a in want.items():
        a64, b64 = np.asarray(a, np.float64), np.asarray(got[k], np.float64)
        scale = np.abs(a64).max()
        ok = bool(np.all(np.abs(a64 - b64) <= 5e-6 * scale + 2e-5 * np.abs(a64)))
        if not ok:
            bad.append("%s: max err %.3e (scale %.3e)" % (k, np.abs(a64 - b64).max(), scale))
    assert not bad, "\n".join(bad)
    fl = s["imove"] == 1
    bd = s["imove"] == -3
    assert (want["shepard"][fl] < 0.999).any() and (np.abs(want["shepard"][bd] - 0.5) < 0.2).any()
    assert np.abs(want["p_bound"][bd]).max() > 0 and np.abs(want["lap_u_bi"][bd]).max() > 0
    assert np.abs(want["u_bounce"] - s["u"]).max() > 0, "the elastic bounce must trigger"


FIELDS = {"r": 1e-6, "u": 2e-5, "rho": 1e-6, "p": 2e-4, "dudt": 2e-4, "drhodt": 5e-4}


def test_dam_break_2d_pipeline(oracle):
    """BASELINE config 1: the unchanged 57-tool pipeline of examples/2D/spheric_testcase5_dambreak
    (improved Euler, delta-SPH full, BI boundaries, elastic bounce, variable time step) on the
    GPU against the oracle interpreter, three steps: neighbour structures and dt bit-exact,
    fields within the stated fp32 tolerance."""
    from aquagpusph_b200 import casegen, host
    from oracle import interp
    host.set_log_level(3)
    case = cases.spheric5_dam_break_2d(3000, 3.0, seed=4)
    xml = casegen.instantiate("spheric5_dambreak_2d", case, (case["N"],))
    I = interp.Interpreter(xml, 2)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = case[k]
    sim = casegen.load("spheric5_dambreak_2d", case, (case["N"],))
    tools = [t for t in sim.tools() if not t[1].startswith("report")]
    assert len(tools) == 57
    assert tools == [(t["name"], t["type"]) for t in I.tools if not t["type"].startswith("report")]
    for step in range(3):
        I.step()
        sim.step(1)
        assert np.array_equal(sim.scalar("n_cells", np.uint32, 4), I.V["n_cells"])
        assert float(sim.scalar("dt")) == float(I.V["dt"]), "dt must be bit-exact"
        if step == 0:
            for k in ("icell", "id_sorted", "id_unsorted"):
                assert np.array_equal(sim.download(k, np.uint32), I.V[k]), k
            ncw = int(I.V["n_cells"][3])
            assert np.array_equal(sim.download("ihoc", np.uint32)[:ncw], I.V["ihoc"][:ncw])
        fl = I.unsorted("imove") == 1
        for k, tol in FIELDS.items():
            a = I.unsorted(k).astype(np.float64)
            b = sim.download(k, unsorted=True).astype(np.float64)
            scale = np.abs(a[fl]).max()
            err = np.abs(a[fl] - b[fl]).max()
            assert err <= tol * scale, "step %d field %s: err %.3e scale %.3e" % (step, k, err, scale)
    assert sim.launch_count() > 0
    sim.close()
