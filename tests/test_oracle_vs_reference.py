"""CPU: pins the oracle (oracle/aqo_*.c, the plain-C restatement) against the
reference's OWN device scripts, compiled as C++ behind oracle/ref_shim from
/root/reference/resources/Scripts (oracle/_ref/libaquaref{2,3}d.so; built in the
build container by __graft_entry__.build()).  Both run the same sequence of
script kernels, arguments bound by name, on the same cell-sorted dam-break
state: fluid, boundary-integral elements, sensors and buffer particles.

Both sides are fp32 without FMA contraction and the restatement keeps the
reference's operation order, so EVERY output must be bit-identical."""
import numpy as np
import pytest

import cases
import pipeline
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available() and not ref.build(),
                                reason="oracle/_ref not built (needs /root/reference once)")

@pytest.mark.parametrize("dims,n,hfac", [(3, 12, 2.0), (2, 50, 3.0), (3, 9, 3.0), (2, 36, 4.0)])
def test_restatement_matches_reference_scripts(oracle, dims, n, hfac):
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    want = pipeline.ref_sweeps(ref.Ref(dims, case["h"]), s)
    got = pipeline.oracle_sweeps(s)
    bad = []
    for k, a in want.items():
        b = got[k]
        a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if not np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True):
            bad.append("%s: max err %.3e (scale %.3e)" % (k, np.abs(a64 - b64).max(), np.abs(a64).max()))
    assert not bad, "\n".join(bad)
    assert np.abs(want["grad_p"]).max() > 0 and np.abs(want["grad_w_bi"]).max() > 0
    assert np.abs(want["lap_p"]).max() > 0 and np.abs(want["dudt"] - want["dudt_pre"]).max() > 0


@pytest.mark.parametrize("dims", [2, 3])
def test_time_schemes_and_permutation_match_reference_scripts(oracle, dims):
    """Element-wise kernels: integrators, Domain, Sort stages, SetBuffer -- bit-exact."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    R = ref.Ref(dims, case["h"])
    rng = np.random.default_rng(5)
    base = {k: np.ascontiguousarray(case[k]).copy() for k in
            ("imove", "iset", "r", "u", "dudt", "rho", "drhodt", "m", "id", "normal", "tangent")}

    def fresh():
        v = {k: a.copy() for k, a in base.items()}
        for k in ("r", "u", "dudt"):
            v[k + "_in"] = rng.normal(size=(N, V)).astype(np.float32)
        for k in ("rho", "drhodt"):
            v[k + "_in"] = rng.normal(size=N).astype(np.float32)
        v.update(N=N, dt=1e-3, relax_midpoint=0.25, domain_min=case["domain_min"],
                 domain_max=case["domain_max"])
        return v

    # midpoint scheme pieces, each against the restatement
    for entry, ofn, outs in (
            ("predictor", lambda v: oracle.call("mp_predictor", v["r"], v["u"], v["dudt"], v["rho"],
                                                v["drhodt"], v["r_in"], v["u_in"], v["dudt_in"],
                                                v["rho_in"], v["drhodt_in"], N, dims),
             ("r_in", "u_in", "dudt_in", "rho_in", "drhodt_in")),
            ("midpoint", lambda v: oracle.call("mp_midpoint", v["imove"], v["u_in"], v["u"], v["dudt"],
                                               v["rho_in"], v["rho"], v["drhodt"], N, 1e-3, dims),
             ("u", "rho")),
            ("relax", lambda v: oracle.call("mp_relax", v["imove"], v["dudt_in"], v["dudt"],
                                            v["drhodt_in"], v["drhodt"], N, 0.25, dims),
             ("dudt", "drhodt")),
            ("corrector", lambda v: oracle.call("mp_corrector", v["imove"], v["r_in"], v["r"], v["u_in"],
                                                v["u"], v["dudt"], v["rho_in"], v["rho"], v["drhodt"],
                                                N, 1e-3, dims),
             ("r", "u", "rho"))):
        rng = np.random.default_rng(5)
        a = fresh()
        rng = np.random.default_rng(5)
        b = fresh()
        R.run("basic/time_scheme/midpoint.cl", entry, N, a)
        ofn(b)
        for k in outs:
            assert np.array_equal(a[k], b[k]), (entry, k)
    # Domain
    rng = np.random.default_rng(5)
    a = fresh()
    a["r_in"][::7] *= 100.0
    a["r_in"][3, 0] = np.nan
    b = {k: (x.copy() if isinstance(x, np.ndarray) else x) for k, x in a.items()}
    R.run("basic/Domain.cl", "entry", N, a)
    oracle.call("domain", b["imove"], b["r_in"], b["u_in"], b["dudt_in"], b["m"], N,
                b["domain_min"], b["domain_max"], dims)
    for k in ("imove", "r_in", "u_in", "dudt_in", "m"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert (a["imove"] == -256).any()
    # basic/Sort.cl stage1 + stage2 == the generic scatter of the restatement
    ll = oracle.linklist(base["r"], dims, 2.0, case["h"])
    v = fresh()
    v["id_sorted"] = ll["inv_perm"]
    for k in ("id", "iset", "imove", "normal", "tangent", "m"):
        v[k + "_in"] = base[k].copy()
    v["r_in"], v["u_in"], v["rho_in"] = base["r"].copy(), base["u"].copy(), base["rho"].copy()
    R.run("basic/Sort.cl", "stage1", N, v)
    R.run("basic/Sort.cl", "stage2", N, v)
    for k in ("id", "iset", "imove", "r", "normal", "tangent", "rho", "m", "u"):
        assert np.array_equal(v[k], oracle.scatter(base[k], ll["inv_perm"])), k
    assert np.array_equal(v["dudt_in"], oracle.scatter(base["dudt"], ll["inv_perm"]))   # Sort.cl:119-123
    assert np.array_equal(v["drhodt_in"], oracle.scatter(base["drhodt"], ll["inv_perm"]))


@pytest.mark.parametrize("dims", [2, 3])
def test_motion_kernels_match_reference_scripts(oracle, dims):
    """cfd/Motions/{Transform,UnTransform,Velocity,Acceleration}.cl (the moving-wall preset,
    cfd/motion.xml): the restatement is bit-identical to the reference's scripts, and Transform
    followed by UnTransform gives the positions back."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    R = ref.Ref(dims, case["h"])
    rng = np.random.default_rng(11)
    base = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "iset", "r", "normal", "tangent")}
    base["iset"] = (np.arange(N) % 2).astype(np.uint32)           # two sets: only set 1 moves
    base["normal"] = rng.normal(size=(N, V)).astype(np.float32)
    base["tangent"] = rng.normal(size=(N, V)).astype(np.float32)
    if dims == 3:
        base["normal"][:, 3] = 0
        base["tangent"][:, 3] = 0
    sc = dict(N=N, motion_iset=1,
              motion_r=np.array([0.3, -0.2, 0.1, 0.0], np.float32)[:V].copy(),
              motion_a=np.array([0.21, -0.13, 0.37, 0.0], np.float32),
              motion_drdt=np.array([0.5, 0.25, -0.125, 0.0], np.float32)[:V].copy(),
              motion_dadt=np.array([0.7, -0.4, 1.1, 0.0], np.float32),
              motion_ddrddt=np.array([-1.5, 0.75, 2.0, 0.0], np.float32)[:V].copy(),
              motion_ddaddt=np.array([0.9, 0.3, -0.6, 0.0], np.float32))
    sc["motion_r_in"], sc["motion_a_in"] = sc["motion_r"], sc["motion_a"]

    def fresh():
        v = {k: a.copy() for k, a in base.items()}
        v["u"] = np.zeros((N, V), np.float32)
        v["dudt"] = np.zeros((N, V), np.float32)
        v.update(sc)
        return v

    a, b = fresh(), fresh()
    moved = (base["iset"] == 1) & (base["imove"] != 1)
    assert moved.any() and (~moved).any()
    # velocity / acceleration are evaluated on the untransformed (local) positions
    R.run("cfd/Motions/Velocity.cl", "entry", N, a)
    oracle.call("motion_velocity", b["iset"], b["imove"], b["r"], b["u"], N, 1, sc["motion_drdt"],
                sc["motion_a"], sc["motion_dadt"], dims)
    R.run("cfd/Motions/Acceleration.cl", "entry", N, a)
    oracle.call("motion_acceleration", b["iset"], b["imove"], b["r"], b["dudt"], N, 1, sc["motion_ddrddt"],
                sc["motion_a"], sc["motion_ddaddt"], dims)
    R.run("cfd/Motions/Transform.cl", "entry", N, a)
    oracle.call("motion_transform", b["iset"], b["imove"], b["r"], b["normal"], b["tangent"], N, 1,
                sc["motion_r"], sc["motion_a"], dims)
    for k in ("u", "dudt", "r", "normal", "tangent"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["r"][~moved], base["r"][~moved]) and np.abs(a["u"][moved]).max() > 0
    assert np.abs(a["r"][moved] - base["r"][moved]).max() > 0.05
    assert np.allclose(np.linalg.norm(a["normal"][moved], axis=1), 1.0, atol=1e-6)
    unit_n = a["normal"].copy()
    R.run("cfd/Motions/UnTransform.cl", "entry", N, a)
    oracle.call("motion_untransform", b["iset"], b["imove"], b["r"], b["normal"], b["tangent"], N, 1,
                sc["motion_r_in"], sc["motion_a_in"], dims)
    for k in ("r", "normal", "tangent"):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(a["r"] - base["r"]).max() < 2e-6 * np.abs(base["r"]).max() + 1e-6   # round trip
    del unit_n


@pytest.mark.parametrize("dims", [2, 3])
def test_energy_kernels_match_reference_scripts(oracle, dims):
    """cfd/Energy/Energy.cl::power and ::energy (preset cfd/energy.xml), bit-identical."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    R = ref.Ref(dims, case["h"])
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
    a = dict(v, N=N, g=g, cs=float(case["cs"]), **{k: np.full(N, 7.0, np.float32) for k in names})
    b = {k: np.full(N, 7.0, np.float32) for k in names}
    R.run("cfd/Energy/Energy.cl", "power", N, a)
    R.run("cfd/Energy/Energy.cl", "energy", N, a)
    oracle.call("energy_power", b["energy_dekdt"], b["energy_depdt"], b["energy_decdt"], v["imove"], v["u"],
                v["rho"], v["m"], v["p"], v["dudt"], v["drhodt"], N, g, dims)
    oracle.call("energy_energy", b["energy_ek"], b["energy_ep"], b["energy_ec"], v["iset"], v["imove"], v["r"],
                v["u"], v["rho"], v["m"], v["refd"], N, g, float(case["cs"]), dims)
    for k in names:
        assert np.array_equal(a[k], b[k]), k
        assert np.abs(a[k]).max() > 0 and (a[k][v["imove"] != 1] == 0).all(), k


@pytest.mark.parametrize("dims", [2, 3])
def test_small_preset_kernels_match_reference_scripts(oracle, dims):
    """cfd/Energy/EnergyKin.cl (cfd/energy_kin.xml), cfd/Forces/Forces.cl (cfd/forces.xml),
    basic/DensityClamp.cl (basic/densityClamp.xml), basic/IdInverse.cl (basic/id_inverse.xml):
    the C restatements are bit-identical to the reference's scripts."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    R = ref.Ref(dims, case["h"])
    rng = np.random.default_rng(8)
    v = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "r", "m", "rho")}
    v["u"] = rng.normal(size=(N, V)).astype(np.float32)
    v["dudt"] = rng.normal(size=(N, V)).astype(np.float32)
    if dims == 3:
        v["u"][:, 3] = 0
        v["dudt"][:, 3] = 0
    g = np.asarray(case["g"], np.float32).ravel()[:V].copy()
    fr = np.array([0.3, -0.1, 0.2, 0.0], np.float32)[:V].copy()
    perm = rng.permutation(N).astype(np.uint32)
    lo, hi = float(np.percentile(v["rho"], 20)), float(np.percentile(v["rho"], 80))
    a = dict(v, N=N, g=g, forces_r=fr, id=perm, rho_min=lo, rho_max=hi, rho_in=v["rho"].copy(),
             energy_kin=np.full(N, 7.0, np.float32), forces_f=np.full((N, V), 7.0, np.float32),
             forces_m=np.full((N, 4), 7.0, np.float32), id_inverse=np.zeros(N, np.uint32))
    b = {k: a[k].copy() for k in ("rho_in", "energy_kin", "forces_f", "forces_m", "id_inverse")}
    R.run("cfd/Energy/EnergyKin.cl", "entry", N, a)
    R.run("cfd/Forces/Forces.cl", "entry", N, a)
    R.run("basic/DensityClamp.cl", "entry", N, a)
    R.run("basic/IdInverse.cl", "entry", N, a)
    oracle.call("energy_kin", b["energy_kin"], v["imove"], v["u"], v["m"], N, dims)
    oracle.call("forces", b["forces_f"], b["forces_m"], v["imove"], v["r"], v["dudt"], v["m"], N, g, fr, dims)
    oracle.call("density_clamp", b["rho_in"], N, lo, hi)
    oracle.call("id_inverse", perm, b["id_inverse"], N)
    for k in b:
        assert a[k].tobytes() == b[k].tobytes(), k
    assert np.array_equal(b["id_inverse"][perm], np.arange(N)) and b["rho_in"].min() == np.float32(lo)
    assert np.abs(b["forces_m"][:, 2]).max() > 0 and (b["forces_f"][v["imove"] != 1] == 0).all()


def _ab_state(case, dims, seed):
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(seed)

    def vec():
        a = rng.normal(size=(N, V)).astype(np.float32)
        if dims == 3:
            a[:, 3] = 0
        return a
    v = {"imove": np.ascontiguousarray(case["imove"]).copy(), "iset": np.zeros(N, np.uint32),
         "r": np.ascontiguousarray(case["r"]).copy(), "u": vec(), "dudt": vec(),
         "rho": np.ascontiguousarray(case["rho"]).copy(), "drhodt": rng.normal(size=N).astype(np.float32),
         "id_sorted": rng.permutation(N).astype(np.uint32)}
    for l in range(1, 5):
        v["dudt_as%d" % l], v["dudt_as%d_in" % l] = vec(), vec()
        v["drhodt_as%d" % l] = rng.normal(size=N).astype(np.float32)
        v["drhodt_as%d_in" % l] = rng.normal(size=N).astype(np.float32)
    return v


def ab_oracle_step(oracle, v, N, dims, dt, it, steps=5):
    """sort -> corrector -> postcorrector of adam_bashforth.cl on the dict v through the C restatement."""
    lv = range(1, 5)
    P = oracle.ptrs
    oracle.call("ab_sort", P(v["dudt_as%d_in" % l] for l in lv), P(v["dudt_as%d" % l] for l in lv),
                P(v["drhodt_as%d_in" % l] for l in lv), P(v["drhodt_as%d" % l] for l in lv), v["id_sorted"], N, dims)
    oracle.call("ab_corrector", v["imove"], v["r"], v["u"], v["dudt"], v["rho"], v["drhodt"],
                P(v["dudt_as%d" % l] for l in lv), P(v["drhodt_as%d" % l] for l in lv), N, float(dt), int(it),
                int(steps), dims)
    oracle.call("ab_postcorrector", P(v["dudt_as%d" % l] for l in lv), P(v["drhodt_as%d" % l] for l in lv),
                v["dudt"], v["drhodt"], P(v["dudt_as%d_in" % l] for l in lv),
                P(v["drhodt_as%d_in" % l] for l in lv), N, dims)


@pytest.mark.parametrize("dims", [2, 3])
def test_adams_bashforth_kernels_match_reference_scripts(oracle, dims):
    """basic/time_scheme/adam_bashforth.cl::sort / ::corrector / ::postcorrector (preset
    basic/time_scheme/adams_bashforth.xml) for iter = 0 .. 6, i.e. every order DYDT_1 .. DYDT_5 of the
    default TSCHEME_ADAMS_BASHFORTH_STEPS = 5: bit-identical.  (::predictor has the body of the midpoint
    predictor, checked in test_time_schemes_and_permutation_match_reference_scripts.)"""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N = case["N"]
    R = ref.Ref(dims, case["h"])
    a = _ab_state(case, dims, 21)
    b = {k: x.copy() for k, x in a.items()}
    dt = 1.25e-3
    for it in range(7):
        A = dict(a, N=N, dt=dt, iter=it)
        R.run("basic/time_scheme/adam_bashforth.cl", "sort", N, A)
        R.run("basic/time_scheme/adam_bashforth.cl", "corrector", N, A)
        R.run("basic/time_scheme/adam_bashforth.cl", "postcorrector", N, A)
        ab_oracle_step(oracle, b, N, dims, dt, it)
        for k in a:
            assert a[k].tobytes() == b[k].tobytes(), (it, k)
        # new rates for the next round
        a["dudt"][:, :dims] = np.random.default_rng(100 + it).normal(size=(N, dims)).astype(np.float32)
        b["dudt"][...] = a["dudt"]
    assert not np.array_equal(a["r"], case["r"])


def noslip_inputs(case, s, seed=4):
    """lap_u as a fluid-fluid sweep would leave it (random here), velocities on fluid and walls."""
    rng = np.random.default_rng(seed)
    N, dims = s["N"], s["dims"]
    V = 4 if dims == 3 else 2
    lap = np.zeros((N, V), np.float32)
    lap[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
    u = np.zeros((N, V), np.float32)
    u[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
    iset = (np.arange(N) % 2).astype(np.uint32)     # only half of the elements belong to the no-slip set
    return lap, u, iset


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0)])
def test_bi_noslip_matches_reference_script(oracle, dims, n, hfac):
    """cfd/Boundary/BI/NoSlip.cl::entry (preset cfd/BINoSlip.xml, the lid-driven cavity of
    examples/2D/spheric_testcase3_liddriven): the C restatement is bit-identical to the script."""
    import pipeline
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    lap, u, iset = noslip_inputs(case, s)
    c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
    c.set("lap_u", lap)
    c.set("u", u)
    c.set("iset", iset)
    c.run("cfd/Boundary/BI/NoSlip.cl", noslip_iset=1, dr=float(case["dr"]))
    got = lap.copy()
    oracle.call("bi_noslip", oracle.make_defs(dims, s["h"]), pipeline._ll(s), iset, s["imove"], s["r"],
                s["normal"], u, s["rho"], s["m"], got, 1, float(case["dr"]))
    assert c.get("lap_u").tobytes() == got.tobytes()
    fl = s["imove"] == 1
    assert np.abs(got - lap)[fl].max() > 0 and np.array_equal(got[~fl], lap[~fl])


@pytest.mark.parametrize("dims", [2, 3])
def test_linklist_tool_kernels_match_the_reference_source(oracle, golden, dims):
    """The tool layer: aquagpusph/CalcServer/LinkList.cl.in (iCell :54-85, iHoc :32-42, linkList
    :92-113), compiled from the reference tree behind the shim with its LinkList.hcl.in, against
    the restatement oracle/aqo_linklist.c -- cell index of every particle and the head-of-cell
    table BIT for bit, on the reference's own LinkList test particles, on a dam break with buffer
    particles far away (a grid much larger than the fluid) and on random positions.  The radix
    sort between the two kernels keeps its property tests (RadixSort.cl.in scans through __local
    memory between barriers: it cannot run one work-item at a time)."""
    R = ref.Ref(dims, 0.1)
    V = 4 if dims == 3 else 2
    rng = np.random.default_rng(17)
    case = cases.dam_break(dims, 12 if dims == 3 else 50, 2.0)
    sets = [(np.ascontiguousarray(golden["linklist_%dD_r" % dims], np.float32), 0.1),
            (case["r"], case["h"])]
    rnd = np.zeros((5000, V), np.float32)
    rnd[:, :dims] = rng.normal(size=(5000, dims)).astype(np.float32) * 3.0
    sets.append((rnd, 0.37))
    for r, h in sets:
        r = np.ascontiguousarray(r, np.float32)
        if r.shape[1] != V:
            rr = np.zeros((r.shape[0], V), np.float32)
            rr[:, :min(V, r.shape[1])] = r[:, :min(V, r.shape[1])]
            r = rr
        N = r.shape[0]
        ll = oracle.linklist(r, dims, 2.0, h)
        # iCell on the unsorted positions == the oracle's sorted cells, un-permuted
        icell = np.zeros(N, np.uint32)
        R.run("LinkList.cl", "iCell", N, dict(icell=icell, r=r, N=N, r_min=ll["rmin"], support=2.0, h=h,
                                               n_cells=ll["ncells"]))
        assert np.array_equal(icell[ll["perm"]], ll["icell"])
        assert np.all(np.diff(ll["icell"].astype(np.int64)) >= 0)
        # iHoc + linkList on the sorted cells
        ncw = int(ll["ncells"][3])
        ihoc = np.zeros(ncw, np.uint32)
        R.run("LinkList.cl", "iHoc", ncw, dict(ihoc=ihoc, N=N, n_cells=ll["ncells"]))
        assert np.all(ihoc == N)
        R.run("LinkList.cl", "linkList", N, dict(icell=ll["icell"], ihoc=ihoc, N=N))
        assert np.array_equal(ihoc, ll["ihoc"][:ncw])


def _symmetry_state(dims, seed=21):
    """A dam break with buffer rows at the end and a symmetry plane through the fluid."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    n, V = case["N"], (4 if dims == 3 else 2)
    nbuf = n            # (room for every particle next to the plane)
    N = n + nbuf
    rng = np.random.default_rng(seed)

    def grow(a, fill):
        out = np.empty((N,) + a.shape[1:], a.dtype)
        out[:n] = a
        out[n:] = fill
        return out
    dmax = np.asarray(case["domain_max"], np.float32).ravel()[:V].copy()
    v = {"imove": grow(np.ascontiguousarray(case["imove"]), -255), "iset": grow(np.ascontiguousarray(case["iset"]).astype(np.int32), 0),
         "r": grow(np.ascontiguousarray(case["r"]), dmax), "m": grow(np.ascontiguousarray(case["m"]), 0),
         "rho": grow(np.ascontiguousarray(case["rho"]), 1000.0)}
    for k in ("normal", "tangent", "u", "dudt"):
        a = rng.normal(size=(N, V)).astype(np.float32)
        if dims == 3:
            a[:, 3] = 0
        v[k] = a
    v["drhodt"] = rng.normal(size=N).astype(np.float32)
    fl = v["imove"] == 1
    x0 = float(np.median(v["r"][fl][:, 0]))
    sr = np.zeros(V, np.float32)
    sr[0] = x0
    sn = np.zeros(V, np.float32)
    sn[0], sn[1] = 0.8, 0.6            # normalised, not axis aligned
    return case, v, N, nbuf, sr, sn, dmax


@pytest.mark.parametrize("dims", [2, 3])
def test_symmetry_mirror_kernels_match_reference_script(oracle, dims):
    """cfd/Boundary/Symmetry/Mirror.cl::detect / feed / set / sort / drop (preset cfd/symmetry.xml, SURVEY
    8(f) row 4) in the order the preset runs them: the C restatement is bit-identical to the
    reference's script, particles next to the plane get a mirrored twin in a buffer row, and drop
    removes what lies beyond the plane."""
    case, v, N, nbuf, sr, sn, dmax = _symmetry_state(dims)
    V = 4 if dims == 3 else 2
    R = ref.Ref(dims, case["h"])
    D = oracle.make_defs(dims, case["h"])
    a = dict(imove=v["imove"].copy(), iset=v["iset"].copy(), r_in=v["r"].copy(), r=v["r"].copy(),
             normal=v["normal"].copy(), tangent=v["tangent"].copy(), m=v["m"].copy(), u_in=v["u"].copy(),
             dudt_in=v["dudt"].copy(), dudt=np.zeros((N, V), np.float32), rho_in=v["rho"].copy(),
             drhodt_in=v["drhodt"].copy(), drhodt=np.zeros(N, np.float32),
             imirror=np.full(N, 7, np.uint32), mirror_src=np.full(N, N, np.uint32),
             mirror_src_in=np.zeros(N, np.uint32), N=N, nbuffer=nbuf, symmetry_r=sr, symmetry_n=sn,
             domain_max=dmax)
    b = {k: (x.copy() if isinstance(x, np.ndarray) else x) for k, x in a.items()}
    # detect
    R.run("cfd/Boundary/Symmetry/Mirror.cl", "detect", N, a)
    oracle.call("sym_detect", D, b["imove"], b["r_in"], b["imirror"], N, sr, sn)
    assert np.array_equal(a["imirror"], b["imirror"]) and 0 < a["imirror"].sum() < (v["imove"] > -255).sum()
    # the radix-sort tool of the preset: keys sorted in place, inverse permutation kept
    perm = np.argsort(a["imirror"], kind="stable").astype(np.uint32)
    inv = np.empty(N, np.uint32)
    inv[perm] = np.arange(N, dtype=np.uint32)
    for d in (a, b):
        d["imirror"] = d["imirror"][perm].copy()
        d["imirror_invperm"] = inv
    R.run("cfd/Boundary/Symmetry/Mirror.cl", "feed", N, a)
    oracle.call("sym_feed", b["imove"], b["iset"], b["imirror"], inv, b["mirror_src"], b["normal"], b["tangent"],
                b["r_in"], N, nbuf, sr, sn, dims)
    R.run("cfd/Boundary/Symmetry/Mirror.cl", "set", N, a)
    oracle.call("sym_set", b["mirror_src"], b["m"], b["u_in"], b["dudt_in"], b["dudt"], b["rho_in"], b["drhodt_in"],
                b["drhodt"], N, sn, dims)
    ids = np.random.default_rng(4).permutation(N).astype(np.uint32)
    for d in (a, b):
        d["mirror_src_in"] = d["mirror_src"].copy()
        d["id_sorted"] = ids
    R.run("cfd/Boundary/Symmetry/Mirror.cl", "sort", N, a)
    oracle.call("sym_sort", b["mirror_src_in"], b["mirror_src"], ids, N)
    for d in (a, b):
        d["r"] = d["r_in"].copy()
    R.run("cfd/Boundary/Symmetry/Mirror.cl", "drop", N, a)
    oracle.call("sym_drop", b["imove"], b["r"], N, sr, sn, dmax, dims)
    for k in ("imove", "iset", "mirror_src", "normal", "tangent", "r_in", "m", "u_in", "dudt_in", "dudt", "rho_in",
              "drhodt_in", "drhodt", "r"):
        assert a[k].tobytes() == b[k].tobytes(), k
    # every detected particle has a twin at its mirror image, the twins lie beyond the plane and are dropped
    src = b["mirror_src_in"]
    twins = np.flatnonzero(src < N)
    assert len(twins) == int(a["imirror"].sum())
    dn = ((b["r_in"][twins] - sr) * sn).sum(1) + ((b["r_in"][src[twins]] - sr) * sn).sum(1)
    assert np.abs(dn).max() < 1e-5
    assert (b["imove"][twins][((b["r_in"][twins] - sr) * sn).sum(1) > 1e-6] == -256).all()


@pytest.mark.parametrize("dims", [2, 3])
def test_ideal_gas_elementwise_kernels_match_reference_scripts(oracle, dims):
    """cfd/ideal_gas/{EOS, Rates, Sort, TimeStep}.cl, riemann/Rates.cl and time_scheme/midpoint.cl
    (predictor, midpoint, relax, corrector): the C restatements are bit-identical to the reference's
    scripts (products, quotients and square roots without contraction; OpenCL's min / max argument order)."""
    a, b, N = _ideal_gas_state(dims, 13)
    R = ref.Ref(dims, a["h"])
    D = oracle.make_defs(dims, a["h"])
    R.run("cfd/ideal_gas/EOS.cl", "entry", N, a)
    oracle.call("ig_eos", b["iset"], b["imove"], b["rho"], b["eint"], b["p"], b["gamma"], N)
    R.run("cfd/ideal_gas/Rates.cl", "entry", N, a)
    oracle.call("ig_rates", b["imove"], b["rho"], b["p"], b["div_u"], b["deintdt"], N)
    R.run("cfd/ideal_gas/TimeStep.cl", "entry", N, a)
    oracle.call("ig_timestep", D, b["dt_var"], b["imove"], b["iset"], b["u"], b["rho"], b["p"], N, a["dt"],
                a["dt_min"], a["courant"], b["div_u"], b["grad_p"], b["gamma"])
    R.run("cfd/ideal_gas/time_scheme/midpoint.cl", "predictor", N, a)
    oracle.call("ig_mp_predictor", b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N)
    R.run("cfd/ideal_gas/riemann/Rates.cl", "entry", N, a)
    oracle.call("ig_riemann_rates", b["imove"], b["work_density"], b["deintdt"], N)
    R.run("cfd/ideal_gas/time_scheme/midpoint.cl", "midpoint", N, a)
    oracle.call("ig_mp_midpoint", b["imove"], b["eint_in"], b["deintdt"], b["eint"], N, a["dt"])
    R.run("cfd/ideal_gas/time_scheme/midpoint.cl", "relax", N, a)
    oracle.call("ig_mp_relax", b["imove"], b["deintdt_in"], b["deintdt"], N, a["relax_midpoint"])
    R.run("cfd/ideal_gas/time_scheme/midpoint.cl", "corrector", N, a)
    oracle.call("ig_mp_corrector", b["imove"], b["eint_in"], b["deintdt"], b["eint"], N, a["dt"])
    # the sort reads the *_in / rate arrays of the step and scatters them (basic/Sort.cl's companion)
    a["eint_in"][...] = a["eint"]
    b["eint_in"][...] = b["eint"]
    R.run("cfd/ideal_gas/Sort.cl", "entry", N, a)
    oracle.call("ig_sort", b["eint_in"], b["eint"], b["deintdt"], b["deintdt_in"], b["id_sorted"], N)
    # the other two time schemes: euler.cl (predictor = the same copies, corrector) and improved_euler.cl
    R.run("cfd/ideal_gas/time_scheme/euler.cl", "predictor", N, a)
    oracle.call("ig_mp_predictor", b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N)
    R.run("cfd/ideal_gas/time_scheme/euler.cl", "corrector", N, a)
    oracle.call("ig_euler_corrector", b["imove"], b["eint"], b["deintdt"], N, a["dt"])
    a["deintdt"][...] = a["work_density"]
    b["deintdt"][...] = b["work_density"]
    R.run("cfd/ideal_gas/time_scheme/improved_euler.cl", "corrector", N, a)
    oracle.call("ig_ie_corrector", b["imove"], b["deintdt"], b["deintdt_in"], b["eint"], N, a["dt"])
    R.run("cfd/ideal_gas/time_scheme/improved_euler.cl", "predictor", N, a)
    oracle.call("ig_ie_predictor", b["imove"], b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N, a["dt"])
    # cfd/ideal_gas/symmetry/Mirror.cl::set (cfd/ideal_gas/symmetry.xml): mirrored rows copy their source's energy
    R.run("cfd/ideal_gas/symmetry/Mirror.cl", "set", N, a)
    oracle.call("ig_sym_set", b["mirror_src"], b["eint_in"], b["deintdt_in"], b["deintdt"], N)
    for k in ("p", "deintdt", "dt_var", "eint", "eint_in", "deintdt_in"):
        assert a[k].tobytes() == b[k].tobytes(), k
    ms = a["mirror_src"] < N
    assert ms.sum() > 10 and np.array_equal(a["eint_in"][ms], a["eint_in"][a["mirror_src"][ms]])
    fl = a["imove"] == 1
    assert np.isfinite(a["dt_var"]).all() and (a["dt_var"][~(a["imove"] > 0)] == np.float32(a["dt"])).all()
    assert (a["dt_var"][fl] < np.float32(a["dt"])).any() and (a["dt_var"] >= np.float32(a["dt_min"])).all()


def _ideal_gas_state(dims, seed):
    """Two equal states (reference / restatement) of a gas: dam-break particle classes, positive rho, p, eint."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(seed)

    def pos(lo, hi):
        return rng.uniform(lo, hi, N).astype(np.float32)

    def vec(s):
        v = (s * rng.normal(size=(N, V))).astype(np.float32)
        if dims == 3:
            v[:, 3] = 0
        return v

    imove = np.ascontiguousarray(case["imove"]).copy()
    imove[rng.random(N) < 0.05] = -1          # (the class EXCLUDED_PARTICLE of the EOS lets through)
    a = dict(imove=imove, iset=(rng.random(N) < 0.5).astype(np.uint32), rho=pos(0.5, 2.0), eint=pos(1.0, 3.0),
             p=pos(0.5, 2.0), div_u=(3 * rng.normal(size=N)).astype(np.float32), deintdt=pos(-1.0, 1.0),
             dt_var=np.zeros(N, np.float32), u=vec(1.0), dudt=vec(1.0), grad_p=vec(5.0), m=pos(0.1, 0.2),
             gamma=np.array([1.4, 1.6667], np.float32), work_density=pos(-1.0, 1.0), eint_in=pos(1.0, 3.0),
             deintdt_in=pos(-1.0, 1.0), id_sorted=rng.permutation(N).astype(np.uint32))
    # mirrored particles (a fifth of the rows) point at sources that are not mirrored themselves
    src = np.full(N, N, np.uint32)
    mirrored = rng.random(N) < 0.2
    src[mirrored] = rng.choice(np.flatnonzero(~mirrored), int(mirrored.sum())).astype(np.uint32)
    a["mirror_src"] = src
    b = {k: v.copy() for k, v in a.items()}
    for d in (a, b):
        d.update(N=N, dt=2.5e-3, dt_min=1e-5, courant=0.25, h=case["h"], relax_midpoint=0.35)
    return a, b, N


def riemann_inputs(s, seed=6):
    """A gas on the sorted dam-break particles: positive pressure, two heat-capacity ratios, random velocities;
    outputs pre-filled (the script overwrites the fluid rows and leaves the others)."""
    rng = np.random.default_rng(seed)
    N, dims = s["N"], s["dims"]
    V = 4 if dims == 3 else 2
    u = np.zeros((N, V), np.float32)
    u[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
    return dict(u=u, p=rng.uniform(0.5, 2.0, N).astype(np.float32), iset=(np.arange(N) % 2).astype(np.uint32),
                gamma=np.array([1.4, 1.6667], np.float32), grad_p=np.full((N, V), 7.0, np.float32),
                div_u=np.full(N, 7.0, np.float32), work_density=np.full(N, 7.0, np.float32))


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0)])
def test_riemann_interactions_match_reference_script(oracle, dims, n, hfac):
    """cfd/ideal_gas/riemann/Interactions.cl::entry (the acoustic Riemann solver between fluid particles,
    examples/2D/shock_point_riemann, shock_1d): the C restatement is bit-identical to the script."""
    import pipeline
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = riemann_inputs(s)
    c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
    for k in ("u", "p", "iset", "grad_p", "div_u"):
        c.set(k, x[k])
    c.v["gamma"] = x["gamma"].copy()
    c.v["work_density"] = x["work_density"].copy()
    c.run("cfd/ideal_gas/riemann/Interactions.cl")
    g, d, w = x["grad_p"].copy(), x["div_u"].copy(), x["work_density"].copy()
    oracle.call("ig_riemann_interactions", oracle.make_defs(dims, s["h"]), pipeline._ll(s), x["iset"], s["imove"],
                s["r"], x["u"], s["rho"], s["m"], x["p"], g, d, w, x["gamma"])
    assert c.get("grad_p").tobytes() == g.tobytes()
    assert c.get("div_u").tobytes() == d.tobytes() and c.get("work_density").tobytes() == w.tobytes()
    fl = s["imove"] == 1
    assert np.isfinite(g).all() and np.abs(g[fl][:, :dims]).max() > 1e-3 and np.abs(w[fl]).max() > 1e-3
    assert (d[~fl] == 7.0).all() and (g[~fl] == 7.0).all()
    if dims == 3:
        assert (g[:, 3] == 7.0).all()     # only .XYZ is written


@pytest.mark.parametrize("dims", [2, 3])
def test_open_boundary_kernels_match_reference_scripts(oracle, dims):
    """cfd/Boundary/Inlet/Inlet.cl::feed / rates, Outlet/Outlet.cl::rates / feed, Portal/Mirror.cl::mirror /
    unmirror / teleport (presets cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml): the C
    restatement gives the reference script's bits after every kernel of the sequence."""
    import open_boundary_common as ob
    case, v = ob.state(dims)
    R = ref.Ref(dims, case["h"])
    D = oracle.make_defs(dims, case["h"])
    a, b = ob.args_of(v), ob.args_of(v)
    for key in ob.STEPS:
        n = v["N"] if key != (ob.INLET, "feed") else v["nbuffer"] + 13   # (more work items than buffer rows)
        R.run(key[0], key[1], n, a)
        ob.oracle_step(oracle, D, dims, key, b)
        for k in ob.WRITES[key]:
            assert a[k].tobytes() == b[k].tobytes(), (key, k)
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert a[k].tobytes() == b[k].tobytes(), k
    ob.checks(v, b, dims)
    # a starving flag of 0 leaves everything alone
    c = ob.args_of(v)
    c["inlet_starving"] = 0
    ob.oracle_step(oracle, D, dims, (ob.INLET, "feed"), c)
    R.run(ob.INLET, "feed", v["nbuffer"], ob.args_of(v), inlet_starving=0)
    assert all(c[k].tobytes() == v[k].tobytes() for k in ob.WRITES[(ob.INLET, "feed")])


def morris_inputs(s, seed=8):
    """Random velocities and pressures on the sorted dam-break particles; outputs pre-filled (the script
    overwrites the XYZ components of the fluid rows and leaves everything else)."""
    rng = np.random.default_rng(seed)
    N, dims = s["N"], s["dims"]
    V = 4 if dims == 3 else 2
    u = np.zeros((N, V), np.float32)
    u[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
    return dict(u=u, p=rng.uniform(-50.0, 300.0, N).astype(np.float32), grad_p=np.full((N, V), 7.0, np.float32),
                lap_u=np.full((N, V), 7.0, np.float32), div_u=np.full(N, 7.0, np.float32))


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0), (2, 36, 4.0)])
def test_interactions_morris_laplacian_matches_reference_script(oracle, dims, n, hfac):
    """cfd/Interactions.cl::entry compiled with __LAP_FORMULATION__ = __LAP_MORRIS__ (the <Define> of
    examples/2D/taylor_green and cylinder_inside_channel; Interactions.cl:130-131): the C restatement is
    bit-identical to the script, and only lap_u differs from the Monaghan build."""
    import pipeline
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = morris_inputs(s)
    c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
    for k in ("u", "p", "grad_p", "lap_u", "div_u"):
        c.set(k, x[k])
    c.run("cfd/Interactions@morris.cl")
    g, l, d = x["grad_p"].copy(), x["lap_u"].copy(), x["div_u"].copy()
    D = oracle.make_defs(dims, s["h"])
    oracle.call("interactions_morris", D, pipeline._ll(s), s["imove"], s["r"], x["u"], s["rho"], s["m"], x["p"],
                g, l, d)
    assert c.get("grad_p").tobytes() == g.tobytes() and c.get("lap_u").tobytes() == l.tobytes()
    assert c.get("div_u").tobytes() == d.tobytes()
    g2, l2, d2 = x["grad_p"].copy(), x["lap_u"].copy(), x["div_u"].copy()
    oracle.call("interactions", D, pipeline._ll(s), s["imove"], s["r"], x["u"], s["rho"], s["m"], x["p"], g2, l2, d2)
    fl = s["imove"] == 1
    assert g2.tobytes() == g.tobytes() and d2.tobytes() == d.tobytes()
    assert np.isfinite(l).all() and np.abs(l[fl][:, :dims] - l2[fl][:, :dims]).max() > 1e-3
    assert (l[~fl] == 7.0).all() and (dims == 2 or (l[:, 3] == 7.0).all())


@pytest.mark.parametrize("morris", [0, 1])
@pytest.mark.parametrize("dims,n,hfac", [(2, 60, 3.0), (3, 24, 1.3), (2, 80, 4.0)])
def test_portal_sweeps_match_reference_scripts(oracle, dims, n, hfac, morris):
    """cfd/Boundary/Portal/Shepard.cl::entry and Portal/Interactions.cl::entry (the latter under both Laplacian
    definitions; preset cfd/portal.xml, examples/2D/taylor_green) on a state Portal/Mirror.cl::mirror has prepared:
    the C restatements are bit-identical to the scripts; only mirrored rows change."""
    import pipeline
    import open_boundary_common as ob
    case, s, D, x = ob.portal_sweep_state(oracle, dims, n, hfac)
    c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
    for k in ("r", "icell", "u", "p", "grad_p", "lap_u", "div_u", "shepard"):
        c.set(k, x[k])
    c.v["imirrored"] = x["imirrored"].copy()
    c.run("cfd/Boundary/Portal/Shepard.cl")
    c.run("cfd/Boundary/Portal/Interactions@morris.cl" if morris else "cfd/Boundary/Portal/Interactions.cl")
    w = ob.portal_sweeps_oracle(oracle, s, D, x, morris)
    for k in w:
        assert c.get(k).tobytes() == w[k].tobytes(), k
    for k in w:
        mir = ob.portal_rows(s, x, k)
        assert np.array_equal(w[k][~mir], x[k][~mir]), k
        assert np.isfinite(w[k]).all() and np.abs(w[k][mir].astype(np.float64) - x[k][mir]).max() > 1e-3, k
