"""GPU parity of the element-wise kernels of the presets next to the hot path (SURVEY 8(f) row 2:
cfd/motion.xml and cfd/energy.xml, the moving-tank case) through the Kernel-tool C-ABI vs the
oracle, which tests/test_oracle_vs_reference.py pins bit-for-bit to the reference's own scripts.

Kept in a file of its own that sorts after the hot-path suites: these kernels were written after
the round's GPU budget was spent, so the first run on a B200 is the driver's; the hot-path parity
tests must not hide behind them under -x."""
import numpy as np
import pytest

import cases
from aquagpusph_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims", [2, 3])
def test_motion_kernels(oracle, dims):
    """cfd/Motions/{Velocity,Acceleration,Transform,UnTransform}.cl (moving walls, preset
    cfd/motion.xml) through the Kernel-tool C-ABI vs the oracle (itself bit-identical to the
    reference's scripts, tests/test_oracle_vs_reference.py).  The launcher evaluates cos / sin of
    the three angles on the host, the kernels are built without FMA contraction: fp32 rounding of a
    handful of products is all that may differ (tolerance 2e-6 of the largest value; in practice the
    arrays are bit-identical)."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(11)
    h = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "iset", "r")}
    h["iset"] = (np.arange(N) % 2).astype(np.uint32)
    h["normal"] = rng.normal(size=(N, V)).astype(np.float32)
    h["tangent"] = rng.normal(size=(N, V)).astype(np.float32)
    if dims == 3:
        h["normal"][:, 3] = 0
        h["tangent"][:, 3] = 0
    h["u"] = np.zeros((N, V), np.float32)
    h["dudt"] = np.zeros((N, V), np.float32)
    sc = dict(N=N, motion_iset=1,
              motion_r=np.array([0.3, -0.2, 0.1, 0.0], np.float32)[:V].copy(),
              motion_a=np.array([0.21, -0.13, 0.37, 0.0], np.float32),
              motion_drdt=np.array([0.5, 0.25, -0.125, 0.0], np.float32)[:V].copy(),
              motion_dadt=np.array([0.7, -0.4, 1.1, 0.0], np.float32),
              motion_ddrddt=np.array([-1.5, 0.75, 2.0, 0.0], np.float32)[:V].copy(),
              motion_ddaddt=np.array([0.9, 0.3, -0.6, 0.0], np.float32))
    sc["motion_r_in"], sc["motion_a_in"] = sc["motion_r"], sc["motion_a"]
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    d = {k: ctx.array(v) for k, v in h.items()}
    d.update(sc)
    o = {k: v.copy() for k, v in h.items()}

    def check(keys, what):
        for k in keys:
            a, b = o[k].astype(np.float64), d[k].get().astype(np.float64)
            assert np.abs(a - b).max() <= 2e-6 * max(np.abs(a).max(), 1.0), (what, k)

    ctx.launch("cfd/Motions/Velocity.cl", "entry", d)
    oracle.call("motion_velocity", o["iset"], o["imove"], o["r"], o["u"], N, 1, sc["motion_drdt"],
                sc["motion_a"], sc["motion_dadt"], dims)
    ctx.launch("cfd/Motions/Acceleration.cl", "entry", d)
    oracle.call("motion_acceleration", o["iset"], o["imove"], o["r"], o["dudt"], N, 1, sc["motion_ddrddt"],
                sc["motion_a"], sc["motion_ddaddt"], dims)
    check(("u", "dudt"), "rates")
    ctx.launch("cfd/Motions/Transform.cl", "entry", d)
    oracle.call("motion_transform", o["iset"], o["imove"], o["r"], o["normal"], o["tangent"], N, 1,
                sc["motion_r"], sc["motion_a"], dims)
    check(("r", "normal", "tangent"), "transform")
    moved = (h["iset"] == 1) & (h["imove"] != 1)
    assert np.abs(d["r"].get()[moved] - h["r"][moved]).max() > 0.05
    assert np.array_equal(d["r"].get()[~moved], h["r"][~moved])
    ctx.launch("cfd/Motions/UnTransform.cl", "entry", d)
    oracle.call("motion_untransform", o["iset"], o["imove"], o["r"], o["normal"], o["tangent"], N, 1,
                sc["motion_r_in"], sc["motion_a_in"], dims)
    check(("r", "normal", "tangent"), "untransform")
    assert np.abs(d["r"].get() - h["r"]).max() < 2e-6 * np.abs(h["r"]).max() + 1e-6   # round trip
    ctx.close()


@pytest.mark.parametrize("dims", [2, 3])
def test_energy_kernels(oracle, dims):
    """cfd/Energy/Energy.cl::power / ::energy (preset cfd/energy.xml) through the C-ABI vs the oracle:
    products and sums without contraction are bit-exact; energy_ec goes through the device's logf
    (tolerance 1e-6 of the largest value: the bracket rho0/rho + log(rho/rho0) - 1 cancels to ~1e-4)."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(3)
    v = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "iset", "r", "rho", "m", "refd")}
    v["u"] = rng.normal(size=(N, V)).astype(np.float32)
    v["dudt"] = rng.normal(size=(N, V)).astype(np.float32)
    if dims == 3:
        v["u"][:, 3] = 0
        v["dudt"][:, 3] = 0
    v["p"] = rng.normal(size=N).astype(np.float32) * 1e3
    v["drhodt"] = rng.normal(size=N).astype(np.float32)
    v["rho"] = (v["rho"] * (1 + 0.01 * rng.normal(size=N))).astype(np.float32)
    g = np.asarray(case["g"], np.float32).ravel()[:V].copy()
    names = ("energy_dekdt", "energy_depdt", "energy_decdt", "energy_ek", "energy_ep", "energy_ec")
    b = {k: np.full(N, 7.0, np.float32) for k in names}
    oracle.call("energy_power", b["energy_dekdt"], b["energy_depdt"], b["energy_decdt"], v["imove"], v["u"],
                v["rho"], v["m"], v["p"], v["dudt"], v["drhodt"], N, g, dims)
    oracle.call("energy_energy", b["energy_ek"], b["energy_ep"], b["energy_ec"], v["iset"], v["imove"], v["r"],
                v["u"], v["rho"], v["m"], v["refd"], N, g, float(case["cs"]), dims)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    d = {k: ctx.array(x) for k, x in v.items()}
    for k in names:
        d[k] = ctx.array(np.full(N, 7.0, np.float32))
    d.update(N=N, g=g, cs=float(case["cs"]))
    ctx.launch("cfd/Energy/Energy.cl", "power", d)
    ctx.launch("cfd/Energy/Energy.cl", "energy", d)
    for k in names:
        got = d[k].get()
        if k == "energy_ec":
            assert np.abs(got.astype(np.float64) - b[k]).max() <= 1e-6 * np.abs(b[k]).max() + \
                2e-7 * float(case["cs"]) ** 2 * np.abs(v["m"]).max(), k
        else:
            assert np.array_equal(got, b[k]), k
    ctx.close()
