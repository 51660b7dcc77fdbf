"""CPU: the ideal-gas family at pipeline level: the circular blast of
examples/2D/shock_point (generator after its Create.py, the unchanged 73-tool pipeline -- midpoint scheme
with autostop / autorelax, the cfd presets, cfd/ideal_gas EOS / Rates / Sort / time_scheme/midpoint -- in the
oracle interpreter, whose ideal-gas kernels are bit-identical to the reference's scripts)."""
import math

import numpy as np

from aquagpusph_b200 import cases, casegen
from oracle import interp


def test_generator_follows_the_example():
    n = 2000
    c = cases.shock_point_2d(n)
    R, R0, gamma = 0.5, 0.2, 1.4
    dr = (math.pi * R ** 2 / n) ** 0.5
    # Create.py:118-133: a square lattice from -R in steps of dr, the corners beyond R left out
    assert abs(c["N"] - n) < 0.03 * n and np.allclose(c["dr"], dr) and np.allclose(c["h"], 2.0 * dr)
    rad = np.sqrt((c["r"].astype(np.float64) ** 2).sum(1))
    assert rad.max() <= R + 1e-6 and rad.max() > R - 1.5 * dr and np.allclose(c["r"][:, 0].min(), -R + dr, atol=0.51 * dr)
    x = np.unique(c["r"][:, 0])
    assert np.allclose(np.diff(x), dr, rtol=1e-4)
    inner = rad < R0
    e1, e2 = 2.0e5 / ((gamma - 1) * 1.00001), 1.0e5 / ((gamma - 1) * 1.00001)
    assert np.allclose(c["eint"][inner], e1) and np.allclose(c["eint"][~inner], e2) and inner.sum() > 0.1 * c["N"]
    assert np.allclose(c["m"], 1.00001 * dr ** 2) and (c["imove"] == 1).all() and (c["u"] == 0).all()
    assert np.allclose(c["cs"], math.sqrt(gamma * 2.0e5 / 1.00001))
    assert np.allclose(c["domain_max"], R + 4 * 2.0 * dr) and np.allclose(c["domain_min"], -(R + 4 * 2.0 * dr))
    assert set(c["placeholders"]) == {"H", "GAMMA", "R"}


def _interpreter(c, overrides=None):
    I = interp.Interpreter(casegen.instantiate("shock_point_2d", c, (c["N"],), overrides), 2)
    for k in casegen.STATE_FIELDS + ("eint", "deintdt"):
        I.V[k][...] = c[k]
    return I


def test_blast_pipeline_in_the_oracle(oracle):
    """The pressure jump at R0 drives the gas outwards, the rim stays where it is (bc.cl), mass and total
    energy are kept, and the midpoint loop stops on its residual."""
    c = cases.shock_point_2d(3000)
    I = _interpreter(c)
    names = [t["name"] for t in I.tools]
    assert len(I.tools) == 71      # 73 minus the two reports instantiate() drops
    for a, b in (("set fixed parts", "predictor"), ("predictor", "predictor ideal gas"),
                 ("predictor ideal gas", "unset fixed parts"), ("sort stage2", "sort ideal gas"),
                 ("sort ideal gas", "EOS"), ("midpoint advance", "midpoint advance ideal gas"),
                 ("cfd rates", "cfd rates ideal gas"), ("midpoint relax", "midpoint relax ideal gas"),
                 ("corrector", "corrector ideal gas")):
        assert names.index(b) == names.index(a) + 1, (a, b)
    scripts = {(t.get("path", "").split("Scripts/")[-1], t.get("entry_point")) for t in I.tools if t["type"] == "kernel"}
    for s in (("cfd/ideal_gas/EOS.cl", "entry"), ("cfd/ideal_gas/Rates.cl", "entry"), ("cfd/ideal_gas/Sort.cl", "entry"),
              ("cfd/ideal_gas/time_scheme/midpoint.cl", "predictor"), ("cfd/ideal_gas/time_scheme/midpoint.cl", "midpoint"),
              ("cfd/ideal_gas/time_scheme/midpoint.cl", "relax"), ("cfd/ideal_gas/time_scheme/midpoint.cl", "corrector"),
              ("bc.cl", "set_fixed"), ("bc.cl", "unset_fixed")):
        assert s in scripts, s
    N, R, R0, h = c["N"], c["R"], c["R0"], c["h"]
    m0 = float(c["m"].astype(np.float64).sum())
    e0 = float((c["m"].astype(np.float64) * c["eint"]).sum())
    rim0 = {}
    ident = c["r"].astype(np.float64)
    for step in range(8):
        I.step()
        assert float(I.V["dt"]) == np.float32(np.float32(0.25) * np.float32(c["h"]) / np.float32(c["cs"]))
        assert 1 <= int(I.V["iter_midpoint"]) <= 11 and (I.V["imove"] == 1).all()
    r = I.unsorted("r").astype(np.float64)
    u = I.unsorted("u").astype(np.float64)
    m = I.unsorted("m").astype(np.float64)
    eint = I.unsorted("eint").astype(np.float64)
    rad0 = np.sqrt((ident ** 2).sum(1))
    # the rim (1.5 kernel supports = 3 h) never moved, the gas next to the jump did, outwards
    rim = rad0 > R - 3.0 * h
    assert rim.sum() > 50 and np.array_equal(r[rim], ident[rim]) and (u[rim] == 0).all()
    ring = np.abs(rad0 - R0) < 1.5 * h
    ur = (u[ring] * ident[ring]).sum(1) / rad0[ring]
    assert ur.mean() > 50.0 and (ur > 0).mean() > 0.95, (ur.mean(), (ur > 0).mean())
    far = (rad0 < 0.3 * R0) | ((rad0 > R0 + 6 * h) & ~rim)
    assert np.abs(u[far]).max() < 1e-2 * np.abs(ur).max()
    # conservation
    assert abs(m.sum() - m0) <= 1e-6 * m0
    etot = float((m * eint).sum() + 0.5 * (m * (u ** 2).sum(1)).sum())
    assert abs(etot - e0) <= 2e-4 * e0, (etot, e0)
    _ = rim0, N
