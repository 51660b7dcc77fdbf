"""BASELINE config 5 on the CPU: the 36-tool lattice pipeline (cases_xml/src/lattice_3d/Main.xml over
the reference's unchanged presets basic + improved_euler + cfd + variableTimeStep) resolves to the
tool list SURVEY 8(d) names, parses identically in the C++ host and the oracle interpreter, names
only registered kernels, and steps in the oracle.  GPU side: tests/test_gpu_presets.py."""
import os
import re

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
