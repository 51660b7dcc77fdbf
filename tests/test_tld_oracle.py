"""BASELINE config 4 on the CPU: the tuned-liquid-damper pipeline (2-D SPHERIC test 9:
examples/2D/spheric_testcase9_tld -- midpoint + BIe + forces + energy + delta-SPH full + motion
presets, 104 tools) resolved by our front-end, its two `python` tools replaced by the prescribed
roll of casegen.prescribed_roll, runs in the oracle interpreter; every script it names is in the
CUDA registry.  The GPU side of the same case is tests/test_gpu_presets.py."""
import re

import numpy as np
import pytest

from aquagpusph_b200 import _lib, casegen, cases


def _xml(case, theta0=0.1, period=0.05, overrides=None):
    txt = casegen.instantiate("spheric9_tld_2d", case, (case["n_set0"], case["n_set1"]), overrides)
    return casegen.prescribed_roll(theta0, period)(txt)


def test_case_generator_follows_the_example():
    """examples/2D/spheric_testcase9_tld/src/Create.py:41-235 at its shipped n = 10000."""
    c = cases.spheric9_tld_2d(10000)
    nx, ny = 313, 32                                   # round(L / dr), round(h / dr) at n = 10000
    assert c["n_set0"] == nx * ny and c["n_set1"] == 2 * nx + 2 * (int(round(0.508 / c["dr"])) + 1)
    assert c["N"] == c["n_set0"] + c["n_set1"]
    f, b = c["imove"] == 1, c["imove"] == -3
    assert f.sum() == c["n_set0"] and b.sum() == c["n_set1"] and (c["iset"][b] == 1).all()
    assert np.allclose(c["m"][b], c["dr"]) and np.allclose(np.linalg.norm(c["normal"][b], axis=1), 1)
    # the fluid sits inside the tank, the elements on its walls, normals pointing outwards
    lo, hi = c["r"][b].min(0), c["r"][b].max(0)
    assert (c["r"][f] > lo).all() and (c["r"][f] < hi).all()
    centre = 0.5 * (lo + hi)
    assert (((c["r"][b] - centre) * c["normal"][b]).sum(1) > 0).all()
    # hydrostatic column: rho = refd + refd g (h - y) / cs^2
    y = c["r"][f][:, 1].astype(np.float64)
    assert np.allclose(c["rho"][f], 998.0 + 998.0 * 9.81 * (ny * c["dr"] - y) / 50.0 ** 2, rtol=1e-6)


def test_every_script_of_the_pipeline_is_in_the_cuda_registry():
    c = cases.spheric9_tld_2d(1500)
    txt = _xml(c)
    assert 'type="python"' not in txt
    tools = re.findall(r'<Tool [^>]*type="kernel"[^>]*path="[^"]*Scripts/([^"]*)" entry_point="([^"]*)"', txt)
    assert len(tools) == 37
    L = _lib.lib()
    for path, entry in set(tools):
        assert L.aqc_kernel_lookup(path.encode(), entry.encode(), 2) >= 0, (path, entry)
    for need in ("cfd/Motions/Transform.cl", "cfd/Energy/Energy.cl", "cfd/Boundary/BIe/PST.cl"):
        assert any(p == need for p, _ in tools), need


def test_host_front_end_and_oracle_agree_on_the_pipeline(tmp_path):
    """The C++ host parses the transformed case (no device needed) into the tool list the oracle
    interpreter runs."""
    from aquagpusph_b200 import host
    from oracle import interp
    c = cases.spheric9_tld_2d(1500)
    txt = _xml(c)
    p = tmp_path / "tld.xml"
    p.write_text(txt)
    tools = host.Simulation(str(p), dims=2, parse_only=True).tools()
    assert tools == [(t["name"], t["type"]) for t in interp.Interpreter(txt, 2).tools]
    names = [n for n, _ in tools]
    # cfd/motion.xml:55-72: after TimeStep, data -> state -> unTransform -> velocity -> acceleration -> transform
    k = names.index("TimeStep")
    assert names[k + 1] == "cfd motion data" and names[k + 4] == "cfd motion state"
    assert names[k + 11:k + 15] == ["cfd motion unTransform", "cfd motion velocity", "cfd motion acceleration",
                                    "cfd motion transform"]


def test_pipeline_runs_in_the_oracle(oracle):
    from oracle import interp
    c = cases.spheric9_tld_2d(1500, 4.0, seed=5)
    I = interp.Interpreter(_xml(c), 2)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    assert len(I.tools) == 108 and I.N == c["N"]          # 104 - 2 python + 3 + 7 set_scalar - 4 reports
    wall = c["imove"] == -3
    w = 2 * np.pi / 0.05
    centre = np.array([0.0, 0.47])
    for step in range(4):
        t = float(I.V["t"])
        I.step()
        th = 0.1 * np.sin(w * t)                          # the tools read t before `t = t + dt`
        assert abs(float(I.V["motion_a"][2]) - th) <= 1e-6 * 0.1
        assert abs(float(I.V["motion_dadt"][2]) - 0.1 * w * np.cos(w * t)) <= 1e-5 * 0.1 * w
        # walls: the file positions rotated by theta around motion_r, rigid-body velocities
        r0 = c["r"][wall].astype(np.float64) - centre
        rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        assert np.abs(I.unsorted("r")[wall] - (r0 @ rot.T + centre)).max() < 2e-6
        om = float(I.V["motion_dadt"][2])
        r1 = I.unsorted("r")[wall].astype(np.float64) - centre
        vel = np.stack([-om * r1[:, 1], om * r1[:, 0]], 1)
        assert np.abs(I.unsorted("u")[wall] - vel).max() < 1e-4 * np.abs(vel).max()
        nrm = np.linalg.norm(I.unsorted("normal")[wall], axis=1)
        assert np.abs(nrm - 1).max() < 1e-6
    # State.py: after the first call the state handed to UnTransform is the previous step's
    assert float(I.V["motion_a_in"][2]) != float(I.V["motion_a"][2]) and int(I.V["motion_first"]) == 0
    # energy report: the kinetic energy integrated from the power equals the one summed directly
    ek, ek_ref = float(I.V["energy_Ek"]), float(I.V["energy_Ek_ref"])
    assert np.isfinite([ek, ek_ref, float(I.V["energy_Ep_ref"]), float(I.V["energy_Ec_ref"])]).all()
    assert float(I.V["energy_Ep_ref"]) > 0 and float(I.V["energy_Ec_ref"]) >= 0
    # the weight of the water column is carried by the bottom: Force_p_y ~ -rho g L h within the
    # discretisation of 1.5 k particles
    weight = 998.0 * 9.81 * 0.9 * 0.092
    assert -1.3 * weight < float(I.V["Force_p"][1]) < -0.6 * weight
    fl = I.unsorted("imove") == 1
    assert np.isfinite(I.unsorted("r")[fl]).all() and np.isfinite(I.unsorted("u")[fl]).all()


def test_missing_python_tool_is_refused():
    """The unchanged template still names the two `python` tools; nothing silently skips them."""
    c = cases.spheric9_tld_2d(1500)
    txt = casegen.instantiate("spheric9_tld_2d", c, (c["n_set0"], c["n_set1"]))
    assert txt.count('type="python"') == 2
    from oracle import interp
    I = interp.Interpreter(txt, 2)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    with pytest.raises(NotImplementedError):
        I.step()
