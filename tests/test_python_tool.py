"""type="python" tools (aquagpusph/CalcServer/Python.cpp:72-205, 295-325): the `aquagpusph` module with
get / set, main() -> bool.  aquagpusph_b200/pytool.py serves the C++ host (through the runner
registered with aqh_set_script_runner) and the oracle interpreter; here it is exercised on the CPU
through the interpreter.  The GPU side is tests/test_gpu_presets.py::test_tld_python_tools."""
import os

import numpy as np
import pytest

from aquagpusph_b200 import casegen, cases, pytool

REF_TLD = "/root/reference/examples/2D/spheric_testcase9_tld/src/templates"


def _interp(txt, case, script_dir, roots=()):
    from oracle import interp
    I = interp.Interpreter(txt, 2)
    I.script_dir, I.script_roots = script_dir, roots
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = case[k]
    return I


def test_python_route_equals_the_set_scalar_route(oracle, tmp_path):
    """cases_xml/scripts/PrescribedRoll.py + MotionState.py through the python tool give, bit for bit,
    what casegen.prescribed_roll's set_scalar tools give (same double-precision arithmetic, narrowed
    once), over four steps of the tuned-liquid-damper pipeline."""
    c = cases.spheric9_tld_2d(1500, 4.0, seed=5)
    base = casegen.instantiate("spheric9_tld_2d", c, (c["n_set0"], c["n_set1"]))
    A = _interp(casegen.python_roll(0.1, 0.05)(base), c, str(tmp_path))
    B = _interp(casegen.prescribed_roll(0.1, 0.05)(base), c, str(tmp_path))
    assert sum(t["type"] == "python" for t in A.tools) == 2 and not any(t["type"] == "python" for t in B.tools)
    for step in range(4):
        A.step()
        B.step()
        for k in ("motion_a", "motion_dadt", "motion_ddaddt", "motion_a_in", "motion_r_in"):
            assert np.array_equal(A.V[k], B.V[k]), (step, k)
        for k in ("r", "u", "dudt", "rho", "normal"):
            assert A.V[k].tobytes() == B.V[k].tobytes(), (step, k)
    assert float(A.V["motion_a"][2]) != 0 and float(A.V["motion_a_in"][2]) != float(A.V["motion_a"][2])


@pytest.mark.skipif(not os.path.isdir(REF_TLD), reason="needs the reference tree (build container only)")
def test_unchanged_reference_scripts_run_in_the_oracle(oracle, tmp_path):
    """BASELINE config 4 as shipped: the UNCHANGED template with the reference's own Motion.py (read
    where it lies) and resources/Scripts/cfd/Motions/State.py.  Motion.py's mechanical model reads
    T_1-94_A100mm_water.dat, which the reference tree does not ship: a synthetic record
    in the same 8-column layout (Motion.py:46-50) -- mass position xi = 0.1 sin(2 pi t / 1.94) -- is
    written next to it."""
    import shutil
    shutil.copy(os.path.join(REF_TLD, "Motion.py"), tmp_path)       # a scratch copy, never committed
    t = np.linspace(0, 2, 2001)
    xi = 0.1 * np.sin(2 * np.pi * t / 1.94)
    dxi = np.gradient(xi, t)
    with open(tmp_path / "T_1-94_A100mm_water.dat", "w") as f:
        f.write("# synthetic record\n# t xi dxi ddxi theta dtheta ddtheta\n")
        for k in range(len(t)):
            f.write(" " + " ".join("%.9g" % v for v in (t[k], xi[k], dxi[k], 0.0, 0.0, 0.0, 0.0)) + "\n")
    c = cases.spheric9_tld_2d(1500, 4.0, seed=5)
    txt = casegen.instantiate("spheric9_tld_2d", c, (c["n_set0"], c["n_set1"]))
    assert txt.count('type="python"') == 2
    I = _interp(txt, c, str(tmp_path), ("/root/reference/resources",))
    a_hist = []
    for step in range(4):
        I.step()
        a_hist.append(float(I.V["motion_a"][2]))
        if step:
            assert float(I.V["motion_a_in"][2]) == a_hist[-2]       # State.py: last step's motion
    # the tank answers the moving mass and the fluid moment: it has started to roll
    assert a_hist[-1] != 0 and abs(a_hist[-1]) < 1e-3 and float(I.V["motion_dadt"][2]) != 0
    rows = (tmp_path / "Motion.dat").read_text().strip().split("\n")
    assert len(rows) == 4 and len(rows[0].split("\t")) == 6           # Motion.py:180-184


class _Vars:
    def __init__(self):
        self.v = {"x": 1.0}

    def py_get(self, name, offset, n):
        return self.v[name]

    def py_set(self, name, value, offset, n):
        self.v[name] = value


def test_script_contract(tmp_path):
    """Python.cpp:304-318: main() must exist, return a bool, and False stops the simulation;
    get / set outside a tool are refused."""
    (tmp_path / "ok.py").write_text("import aquagpusph as aqua\ndef main():\n"
                                    "    aqua.set('x', aqua.get('x') + 1)\n    return True\n")
    (tmp_path / "stop.py").write_text("def main():\n    return False\n")
    (tmp_path / "none.py").write_text("def main():\n    pass\n")
    (tmp_path / "nomain.py").write_text("x = 1\n")
    vars_ = _Vars()
    R = pytool.ScriptRunner(vars_, str(tmp_path))
    R.run("ok.py")
    R.run("ok")                 # by module name, as the reference imports it
    assert vars_.v["x"] == 3.0
    with pytest.raises(pytool.ScriptError, match="simulation stop"):
        R.run("stop.py")
    with pytest.raises(pytool.ScriptError, match="non boolean"):
        R.run("none.py")
    with pytest.raises(pytool.ScriptError, match="main"):
        R.run("nomain.py")
    with pytest.raises(pytool.ScriptError, match="cannot be imported"):
        R.run("missing.py")
    import aquagpusph
    with pytest.raises(RuntimeError, match="outside"):
        aquagpusph.get("x")
    assert pytool.narrow(np.float64(0.1), np.float32, 1) == np.float32(0.1)
    with pytest.raises(ValueError):
        pytool.narrow(np.zeros(3), np.float32, 4)


def test_type_info():
    from aquagpusph_b200 import host
    assert host.type_info("vec", 2) == (np.float32, 2) and host.type_info("vec*", 3) == (np.float32, 4)
    assert host.type_info("vec4", 2) == (np.float32, 4) and host.type_info("unsigned int", 3) == (np.uint32, 1)
    assert host.type_info("matrix*", 2) == (np.float32, 4) and host.type_info("svec4", 3) == (np.uint32, 4)
    with pytest.raises(host.HostError):
        host.type_info("double", 3)


def test_python_command_line(tmp_path, capsys):
    """python -m aquagpusph_b200: the CLI's flags through the binding (so python tools have a runner).
    --resolve needs no device; a run without one fails loudly, never on a CPU path."""
    from aquagpusph_b200 import __main__ as cli
    c = cases.spheric9_tld_2d(1500)
    xml = tmp_path / "Main.xml"
    xml.write_text(casegen.python_roll()(casegen.instantiate("spheric9_tld_2d", c, (c["n_set0"], c["n_set1"]))))
    out = tmp_path / "flat.xml"
    assert cli.main(["-i", str(xml), "-d", "2", "--resolve", str(out), "-l", "3"]) == 0
    assert out.read_text().count('type="python"') == 2
    import torch
    if not torch.cuda.is_available():
        assert cli.main(["-i", str(xml), "-d", "2", "--steps", "1", "-l", "3"]) == 1
        assert "no CPU fallback" in capsys.readouterr().err
