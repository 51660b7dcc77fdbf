"""The lid-driven cavity (SPHERIC test 3, examples/2D/spheric_testcase3_liddriven: improved Euler,
delta-SPH full, BI boundary integrals, BINoSlip, elastic bounce; 55 tools) on the CPU: generator after
the example's Create.py, every script in the CUDA registry, host front-end == oracle interpreter, and
the physics of the first steps in the oracle.  GPU side: tests/test_gpu_presets.py."""
import re

import numpy as np

from aquagpusph_b200 import _lib, casegen, cases, host


def test_case_generator_follows_the_example():
    """Create.py:41-230 at its shipped nx = ny = 200."""
    c = cases.spheric3_lid_driven_2d(200)
    assert c["n_set0"] == 200 * 200 and c["n_set1"] == 4 * 200 and abs(c["dr"] - 1 / 200) < 1e-9
    f, b = c["imove"] == 1, c["imove"] == -3
    assert np.allclose(c["m"][f], 1.0 / 200 ** 2) and np.allclose(c["m"][b], 1 / 200)
    assert c["p0"] == 3.0 and float(c["visc_dyn"][0]) == np.float32(1e-3) and float(c["delta"][0]) == 1.0
    lid = b & (c["r"][:, 1] == 0.5)
    assert lid.sum() == 200 and (c["u"][lid, 0] == 1).all() and (c["u"][~lid] == 0).all()
    assert (np.abs(c["r"][b]).max(1) == 0.5).all() and (np.abs(c["r"][f]) < 0.5).all()
    assert (((c["r"][b]) * c["normal"][b]).sum(1) > 0).all()          # outward normals


def test_pipeline(oracle, tmp_path):
    from oracle import interp
    c = cases.spheric3_lid_driven_2d(40)
    txt = casegen.instantiate("spheric3_liddriven_2d", c, (c["n_set0"], c["n_set1"]))
    p = tmp_path / "cavity.xml"
    p.write_text(txt)
    tools = host.Simulation(str(p), dims=2, parse_only=True).tools()
    I = interp.Interpreter(txt, 2)
    assert tools == [(t["name"], t["type"]) for t in I.tools] and len(tools) == 55
    names = [n for n, _ in tools]
    assert names.index("cfd BI interactions") < names.index("cfd BI no-slip") < names.index("cfd rates")
    L = _lib.lib()
    for path, entry in set(re.findall(r'type="kernel"[^>]*path="[^"]*Scripts/([^"]*)" entry_point="([^"]*)"', txt)):
        assert L.aqc_kernel_lookup(path.encode(), entry.encode(), 2) >= 0, (path, entry)
    assert int(I.V["noslip_iset"]) == 1 and float(I.V["p0"]) == 3.0
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    nf = c["n_fluid"]
    y = c["r"][:nf, 1]
    near_lid = y > 0.5 - 2 * c["dr"]
    drag = []
    for _ in range(4):
        I.step()
        u = I.unsorted("u")[:nf]
        drag.append(float(u[near_lid, 0].mean()))
        assert np.isfinite(u).all()
    # the no-slip lid drags the layers next to it along +x, more every step; the bulk barely moves
    assert drag[0] == 0 and 0 < drag[1] < drag[2] < drag[3]
    assert abs(float(u[~near_lid, 0].mean())) < 0.05 * drag[3]
    # without the no-slip sweep the walls are free-slip: nothing drags the layers next to the lid
    J = interp.Interpreter(re.sub(r'\s*<Tool [^>]*name="cfd BI no-slip"[^>]*/>', "", txt), 2)
    for k in casegen.STATE_FIELDS:
        J.V[k][...] = c[k]
    for _ in range(4):
        J.step()
    assert abs(float(J.unsorted("u")[:nf][near_lid, 0].mean())) < 0.02 * drag[3]
