"""GPU parity of the WHOLE per-step pipeline: the C++ host (XML front-end +
scheduler + libaquacuda kernels) against the independent Python/C oracle
interpreter, both driven by the same resolved XML of the reference's 3-D
dam-break example (116 tools, midpoint inner loop, delta-SPH, BIe boundaries,
variable time step).

Bar: neighbour structures (icell, ihoc, id_sorted, id_unsorted, n_cells) bit-exact
on identical inputs; the time step dt bit-exact (min is order independent);
fields within |gpu - cpu| <= 1e-5 * max|cpu| after N steps, dudt/drhodt looser
(differences of large terms), as SURVEY section 7 states."""
import numpy as np
import pytest

from aquagpusph_b200 import cases, casegen, host

pytestmark = pytest.mark.gpu

FIELDS = {"r": 1e-6, "u": 1e-5, "rho": 1e-6, "p": 2e-4, "dudt": 2e-4, "drhodt": 2e-4}


def _oracle(case, overrides, nset):
    from oracle import interp
    xml = casegen.instantiate("spheric2_dambreak_3d", case, nset, overrides)
    I = interp.Interpreter(xml, 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = case[k]
    return I


@pytest.mark.parametrize("n,maxiter,fusion", [(6000, 30, True), (12000, 3, True), (6000, 3, False)])
def test_dam_break_3d_steps(oracle, n, maxiter, fusion, monkeypatch):
    host.set_log_level(3)
    if not fusion:
        monkeypatch.setenv("AQUA_NO_FUSION", "1")
    ov = {"iter_midpoint_max": maxiter}
    case = cases.spheric2_dam_break(n, 3.0, seed=7)
    nset = (case["N"] - 8, 8)
    I = _oracle(case, ov, nset)
    sim = casegen.load("spheric2_dambreak_3d", case, nset, ov)
    assert [t for t in sim.tools() if not t[1].startswith("report")] == \
        [(t["name"], t["type"]) for t in I.tools if not t["type"].startswith("report")]
    # the four fluid-fluid sweeps of the inner loop run as one fused launch (or not at all)
    assert sim.fused_groups() == (1 if fusion else 0)
    for step in range(3):
        I.step()
        sim.step(1)
        # scalars driven by device results
        assert int(sim.scalar("iter_midpoint", np.uint32)) == int(I.V["iter_midpoint"])
        assert np.array_equal(sim.scalar("n_cells", np.uint32, 4), I.V["n_cells"])
        assert float(sim.scalar("dt")) == float(I.V["dt"]), "dt must be bit-exact"
        if step == 0:
            # identical inputs -> identical neighbour structures
            for k in ("icell", "id_sorted", "id_unsorted"):
                assert np.array_equal(sim.download(k, np.uint32), I.V[k]), k
            ncw = int(I.V["n_cells"][3])
            assert np.array_equal(sim.download("ihoc", np.uint32)[:ncw], I.V["ihoc"][:ncw])
            assert np.array_equal(sim.download("imove", np.int32), I.V["imove"])
        rel = float(sim.scalar("Residual_midpoint")) / max(float(I.V["Residual_midpoint"]), 1e-30)
        assert step == 0 or abs(rel - 1.0) < 1e-3
        for k, tol in FIELDS.items():
            a = I.unsorted(k).astype(np.float64)
            b = sim.download(k, unsorted=True).astype(np.float64)
            fl = I.unsorted("imove") == 1
            scale = np.abs(a[fl]).max()
            err = np.abs(a[fl] - b[fl]).max()
            assert err <= tol * scale, "step %d field %s: err %.3e scale %.3e" % (step, k, err, scale)
    sim.close()


def test_host_errors_are_loud():
    """Unknown scripts / bad expressions fail at setup like the reference does."""
    case = cases.spheric2_dam_break(3000, 3.0)
    nset = (case["N"] - 8, 8)
    txt = casegen.instantiate("spheric2_dambreak_3d", case, nset)
    import os, tempfile
    d = tempfile.mkdtemp()
    bad = txt.replace("Scripts/cfd/Interactions.cl", "Scripts/cfd/NoSuchScript.cl")
    p = os.path.join(d, "bad.xml")
    open(p, "w").write(bad)
    with pytest.raises(host.HostError, match="not in the CUDA kernel registry"):
        host.Simulation(p, dims=3, device=0)
    bad = txt.replace('condition="dt &gt; 0.0"', 'condition="dt &gt; nonexistent_var"')
    open(p, "w").write(bad)
    with pytest.raises(host.HostError):
        host.Simulation(p, dims=3, device=0)
