"""CPU check of the LOGIC of a sweep policy written without a GPU at hand (PBINoSlip, sweeps.cu): the
policy struct and the helpers it uses (ldvec / stvec_xyz / dist2 / q_of / Wend / PBase) are lifted out
of sweeps.cu and sweep.cuh as text, compiled for the host by g++ behind a shim (__ldg = a plain load,
the two MUFU approximations = sqrtf and 1/x), and driven by a brute-force loop with the engines'
contract (sweep.cuh: load_i; for every j: stage_j, skip dead rows, test, body; store_i) instead of the
cell walk.  The engines themselves are verified on the GPU through the other policies; what this pins
is the policy: which pairs count, the formula, what is stored.  Compared with the oracle (bit-identical
to the reference's script) at the sweep tests' tolerance; tests/test_gpu_presets.py::test_bi_noslip_sweep
is the run on a B200."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aquagpusph_b200", "csrc")

SHIM = r"""
#include <math.h>
#include <stddef.h>
#include <stdint.h>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
#define __device__
#define __forceinline__ inline
#define __restrict__
template <class T> static inline T __ldg(const T* p) { return *p; }
"""

DRIVER = r"""
template <int D> static void run(const uint32_t* iset, const int* imove, const void* r, const void* normal,
                                 const void* u, const float* rho, const float* m, void* lap_u, uint32_t N,
                                 uint32_t noslip_iset, float dr, float H, float CONW, float SUPPORT)
{
    PBINoSlip<D> p;
    p.imove = imove; p.invH = 1.f / H; p.cut2 = (SUPPORT * H) * (SUPPORT * H);      // set_base
    p.iset = iset; p.r = r; p.normal = normal; p.u = u; p.rho = rho; p.m = m; p.lap_u = lap_u;
    p.noslip_iset = noslip_iset; p.dr = dr; p.cW = Wend<D>::W * CONW; p.H2 = H * H;   // run_bi_noslip
    for (uint32_t i = 0; i < N; i++) {
        if (!p.i_active(imove[i]))
            continue;
        typename PBINoSlip<D>::IState s;
        p.load_i(s, i);
        for (uint32_t j = 0; j < N; j++) {
            float4 row[PBINoSlip<D>::NJ4];
            p.stage_j(j, row);
            if (!PBase::j_live(row[0]) || !p.test(s, row[0]))
                continue;
            p.body(s, row, 1);
        }
        p.store_i(s, i);
    }
}
template <class P, class Set> static void sweep_all(P& p, const int* imove, uint32_t N, float H, float SUPPORT, Set set)
{
    p.imove = imove; p.invH = 1.f / H; p.cut2 = (SUPPORT * H) * (SUPPORT * H);      // set_base
    set(p);
    for (uint32_t i = 0; i < N; i++) {
        if (!p.i_active(imove[i]))
            continue;
        typename P::IState s;
        p.load_i(s, i);
        for (uint32_t j = 0; j < N; j++) {
            float4 row[P::NJ4];
            p.stage_j(j, row);
            if (!PBase::j_live(row[0]) || !p.test(s, row[0]))
                continue;
            p.body(s, row, 1);
        }
        p.store_i(s, i);
    }
}
template <int D> static void run_inter(const int* imove, const void* r, const void* normal, const void* u,
                                       const float* rho, const float* m, const float* pp, void* grad_p,
                                       float* div_u, uint32_t N, float H, float CONW, float SUPPORT)
{
    PBIInteractions<D> p;
    sweep_all(p, imove, N, H, SUPPORT, [&](PBIInteractions<D>& q) {      // run_bi_inter
        q.r = r; q.normal = normal; q.u = u; q.rho = rho; q.m = m; q.p = pp; q.grad_p = grad_p; q.div_u = div_u;
        q.cW = Wend<D>::W * CONW;
    });
}
extern "C" void emu_bi_inter(int dims, const int* imove, const void* r, const void* normal, const void* u,
                             const float* rho, const float* m, const float* pp, void* grad_p, float* div_u,
                             uint32_t N, float H, float CONW, float SUPPORT)
{
    if (dims == 3)
        run_inter<3>(imove, r, normal, u, rho, m, pp, grad_p, div_u, N, H, CONW, SUPPORT);
    else
        run_inter<2>(imove, r, normal, u, rho, m, pp, grad_p, div_u, N, H, CONW, SUPPORT);
}
template <int D> static void run_riemann(const uint32_t* iset, const int* imove, const void* r, const void* u,
                                         const float* rho, const float* m, const float* pp, void* grad_p, float* div_u,
                                         float* work_density, const float* gamma, uint32_t N, float H, float CONW,
                                         float SUPPORT)
{
    PRiemann<D> p;
    sweep_all(p, imove, N, H, SUPPORT, [&](PRiemann<D>& q) {      // run_ig_riemann
        q.iset = iset; q.r = r; q.u = u; q.rho = rho; q.m = m; q.p = pp; q.grad_p = grad_p; q.div_u = div_u;
        q.work_density = work_density; q.gamma = gamma;
        q.cF = 2.f / H * (Wend<D>::F * CONW);
    });
}
extern "C" void emu_riemann(int dims, const uint32_t* iset, const int* imove, const void* r, const void* u,
                            const float* rho, const float* m, const float* pp, void* grad_p, float* div_u,
                            float* work_density, const float* gamma, uint32_t N, float H, float CONW, float SUPPORT)
{
    if (dims == 3)
        run_riemann<3>(iset, imove, r, u, rho, m, pp, grad_p, div_u, work_density, gamma, N, H, CONW, SUPPORT);
    else
        run_riemann<2>(iset, imove, r, u, rho, m, pp, grad_p, div_u, work_density, gamma, N, H, CONW, SUPPORT);
}
template <int D, class P> static void run_fluid(const int* imove, const void* r, const void* u, const float* rho,
                                                 const float* m, const float* pp, void* grad_p, void* lap_u,
                                                 float* div_u, uint32_t N, float H, float CONF, float SUPPORT)
{
    P p;
    sweep_all(p, imove, N, H, SUPPORT, [&](P& q) {      // run_interactions
        q.r = r; q.u = u; q.rho = rho; q.m = m; q.p = pp; q.grad_p = grad_p; q.lap_u = lap_u; q.div_u = div_u;
        q.cF = Wend<D>::F * CONF;
        q.eps2 = 0.01f * H * H;
    });
}
extern "C" void emu_interactions(int dims, int morris, const int* imove, const void* r, const void* u,
                                 const float* rho, const float* m, const float* pp, void* grad_p, void* lap_u,
                                 float* div_u, uint32_t N, float H, float CONF, float SUPPORT)
{
    if (dims == 3 && morris)
        run_fluid<3, PInteractionsMorris<3>>(imove, r, u, rho, m, pp, grad_p, lap_u, div_u, N, H, CONF, SUPPORT);
    else if (dims == 3)
        run_fluid<3, PInteractions<3>>(imove, r, u, rho, m, pp, grad_p, lap_u, div_u, N, H, CONF, SUPPORT);
    else if (morris)
        run_fluid<2, PInteractionsMorris<2>>(imove, r, u, rho, m, pp, grad_p, lap_u, div_u, N, H, CONF, SUPPORT);
    else
        run_fluid<2, PInteractions<2>>(imove, r, u, rho, m, pp, grad_p, lap_u, div_u, N, H, CONF, SUPPORT);
}
template <int D, class P, class Set> static void sweep_portal(P& p, const int* imirrored, uint32_t N, float H,
                                                              float SUPPORT, Set set)
{
    sweep_all(p, imirrored, N, H, SUPPORT, set);      // set_base(p, ctx, imirrored): the i filter reads imirrored
}
extern "C" void emu_portal(int dims, int morris, const int* imove, const int* imirrored, const void* r, const void* u,
                           const float* rho, const float* m, const float* pp, void* grad_p, void* lap_u, float* div_u,
                           float* shepard, uint32_t N, float H, float CONW, float CONF, float SUPPORT)
{
#define PORTAL(D, MO)                                                                                       \
    {                                                                                                       \
        PPortalShepard<D> a;                                                                                \
        sweep_portal<D>(a, imirrored, N, H, SUPPORT, [&](PPortalShepard<D>& q) {                            \
            q.mv = imove; q.r = r; q.rho = rho; q.m = m; q.shepard = shepard; q.cW = Wend<D>::W * CONW;     \
        });                                                                                                 \
        PPortalInteractions<D, MO> b;                                                                       \
        sweep_portal<D>(b, imirrored, N, H, SUPPORT, [&](PPortalInteractions<D, MO>& q) {                   \
            q.mv = imove; q.r = r; q.u = u; q.rho = rho; q.m = m; q.p = pp; q.grad_p = grad_p;              \
            q.lap_u = lap_u; q.div_u = div_u; q.cF = Wend<D>::F * CONF; q.eps2 = 0.01f * H * H;             \
        });                                                                                                 \
    }
    if (dims == 3 && morris) PORTAL(3, true)
    else if (dims == 3) PORTAL(3, false)
    else if (morris) PORTAL(2, true)
    else PORTAL(2, false)
#undef PORTAL
}
extern "C" void emu_noslip(int dims, const uint32_t* iset, const int* imove, const void* r, const void* normal,
                           const void* u, const float* rho, const float* m, void* lap_u, uint32_t N,
                           uint32_t noslip_iset, float dr, float H, float CONW, float SUPPORT)
{
    if (dims == 3)
        run<3>(iset, imove, r, normal, u, rho, m, lap_u, N, noslip_iset, dr, H, CONW, SUPPORT);
    else
        run<2>(iset, imove, r, normal, u, rho, m, lap_u, N, noslip_iset, dr, H, CONW, SUPPORT);
}
"""


def _between(src, a, b):
    i = src.index(a)
    return src[i:src.index(b, i)]


def _lift():
    cuh = open(os.path.join(CSRC, "sweep.cuh")).read()
    cu = open(os.path.join(CSRC, "sweeps.cu")).read()
    far = re.search(r"constexpr float AQC_FAR = [^;]*;", cuh).group(0)
    dist2 = _between(cuh, "template <int DIMS>\n__device__ __forceinline__ float dist2", "// Policy concept")
    helpers = _between(cu, "constexpr float iM_PI", "// ------------------------------------------------------------------------\n// cfd/Interactions.cl")
    # the two MUFU approximations (1 ulp) become their exact counterparts on the host
    helpers = re.sub(r'asm\("sqrt\.approx\.ftz\.f32[^\n]*\n', "r = sqrtf(x);\n", helpers)
    helpers = re.sub(r'asm\("rcp\.approx\.ftz\.f32[^\n]*\n', "r = 1.f / x;\n", helpers)
    assert "asm(" not in helpers and "struct PBase" in helpers
    policy = _between(cu, "template <int D>\nstruct PBINoSlip : PBase {", "// cfd/Boundary/ElasticBounce.cl:77-148")
    # a policy that IS verified on the GPU (tests/test_gpu_bi.py), to validate this harness itself
    known = _between(cu, "template <int D>\nstruct PBIInteractions : PBase {", "// BI/NoSlip.cl:52-130")
    fluid = _between(cu, "template <int D>\nstruct PInteractions : PBase {",
                     "// ------------------------------------------------------------------------\n// basic/Shepard.cl")
    assert "struct PInteractionsMorris : PInteractions<D>" in fluid
    fluid += _between(cu, "template <int D>\nstruct PPortalShepard : PBase {",
                      "// ------------------------------------------------------------------------\n// basic/deltaSPH.cl")
    return far + "\n" + dist2 + helpers + known + policy + fluid


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_sweep")
    cpp, so = str(d / "emu.cpp"), str(d / "libemu.so")
    open(cpp, "w").write(SHIM + _lift() + DRIVER)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-o", so, cpp])
    return C.CDLL(so)


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0), (2, 40, 4.0)])
def test_noslip_policy_matches_the_oracle(oracle, emu, dims, n, hfac):
    from test_oracle_vs_reference import noslip_inputs
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    lap, u, iset = noslip_inputs(case, s)
    D = oracle.make_defs(dims, s["h"])
    want = lap.copy()
    oracle.call("bi_noslip", D, pipeline._ll(s), iset, s["imove"], s["r"], s["normal"], u, s["rho"], s["m"], want,
                1, float(case["dr"]))
    got = lap.copy()
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = {k: np.ascontiguousarray(s[k]) for k in ("imove", "r", "normal", "rho", "m")}
    emu.emu_noslip(dims, P(iset), P(arr["imove"]), P(arr["r"]), P(arr["normal"]), P(u), P(arr["rho"]), P(arr["m"]),
                   P(got), s["N"], 1, C.c_float(case["dr"]), C.c_float(D.H), C.c_float(D.CONW), C.c_float(D.SUPPORT))
    a, b = want.astype(np.float64), got.astype(np.float64)
    assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a).max() + 2e-5 * np.abs(a)), np.abs(a - b).max()
    fl = s["imove"] == 1
    assert np.abs(want - lap)[fl].max() > 1e-3 and np.array_equal(got[~fl], lap[~fl])


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0)])
def test_harness_reproduces_a_gpu_verified_policy(oracle, emu, dims, n, hfac):
    """The harness itself: PBIInteractions (cfd/Boundary/BI/Interactions.cl), whose GPU parity is established
    (tests/test_gpu_bi.py), lifted and driven the same way, equals the reference's own script (behind the
    shim) at that test's tolerance."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    rng = np.random.default_rng(9)
    N, V = s["N"], (4 if dims == 3 else 2)
    c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
    c.run("basic/EOS.cl")
    gp = np.zeros((N, V), np.float32)
    gp[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
    du = rng.normal(size=N).astype(np.float32)
    c.set("grad_p", gp)
    c.set("div_u", du)
    p_arr, rho = c.get("p"), c.get("rho")
    c.run("cfd/Boundary/BI/Interactions.cl")
    want_g, want_d = c.get("grad_p"), c.get("div_u")
    got_g, got_d = gp.copy(), du.copy()
    D = oracle.make_defs(dims, s["h"])
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = {k: np.ascontiguousarray(s[k]) for k in ("imove", "r", "normal", "u", "m")}
    emu.emu_bi_inter(dims, P(arr["imove"]), P(arr["r"]), P(arr["normal"]), P(arr["u"]), P(rho), P(arr["m"]), P(p_arr),
                     P(got_g), P(got_d), N, C.c_float(D.H), C.c_float(D.CONW), C.c_float(D.SUPPORT))
    for a, b in ((want_g, got_g), (want_d, got_d)):
        a, b = a.astype(np.float64), b.astype(np.float64)
        assert np.all(np.abs(a - b) <= 5e-6 * np.abs(a).max() + 2e-5 * np.abs(a)), np.abs(a - b).max()
    assert np.abs(want_g - gp).max() > 0


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0), (2, 40, 4.0)])
def test_riemann_policy_matches_the_oracle(oracle, emu, dims, n, hfac):
    """PRiemann (cfd/ideal_gas/riemann/Interactions.cl): the policy, driven by the brute-force pair loop, against
    the oracle (bit-identical to the script, tests/test_oracle_vs_reference.py) at the sweep tests' tolerance
    -- the policy folds constants and divides by rho_i once per particle instead of once per pair."""
    from test_oracle_vs_reference import riemann_inputs
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = riemann_inputs(s)
    D = oracle.make_defs(dims, s["h"])
    want = {k: x[k].copy() for k in ("grad_p", "div_u", "work_density")}
    oracle.call("ig_riemann_interactions", D, pipeline._ll(s), x["iset"], s["imove"], s["r"], x["u"], s["rho"],
                s["m"], x["p"], want["grad_p"], want["div_u"], want["work_density"], x["gamma"])
    got = {k: x[k].copy() for k in want}
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = {k: np.ascontiguousarray(s[k]) for k in ("imove", "r", "rho", "m")}
    emu.emu_riemann(dims, P(x["iset"]), P(arr["imove"]), P(arr["r"]), P(x["u"]), P(arr["rho"]), P(arr["m"]), P(x["p"]),
                    P(got["grad_p"]), P(got["div_u"]), P(got["work_density"]), P(x["gamma"]), s["N"],
                    C.c_float(D.H), C.c_float(D.CONW), C.c_float(D.SUPPORT))
    fl = s["imove"] == 1
    for k in want:
        a, b = want[k].astype(np.float64), got[k].astype(np.float64)
        assert np.isfinite(b).all(), k
        assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a[fl]).max() + 2e-5 * np.abs(a)), (k, np.abs(a - b).max())
        assert np.array_equal(got[k][~fl], x[k][~fl]) and np.abs(want[k][fl] - x[k][fl]).max() > 1e-3, k


@pytest.mark.parametrize("morris", [0, 1])
@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0), (2, 40, 4.0)])
def test_interactions_policies_match_the_oracle(oracle, emu, dims, n, hfac, morris):
    """PInteractions (GPU-verified: tests/test_gpu_kernels.py) and PInteractionsMorris (cfd/Interactions.cl under
    __LAP_FORMULATION__ = __LAP_MORRIS__, written without a GPU at hand), driven by the brute-force pair loop,
    against the oracle (both bit-identical to the reference's script under the respective definition,
    tests/test_oracle_vs_reference.py) at the sweep tests' tolerance."""
    from test_oracle_vs_reference import morris_inputs
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = morris_inputs(s)
    D = oracle.make_defs(dims, s["h"])
    want = {k: x[k].copy() for k in ("grad_p", "lap_u", "div_u")}
    oracle.call("interactions_morris" if morris else "interactions", D, pipeline._ll(s), s["imove"], s["r"], x["u"],
                s["rho"], s["m"], x["p"], want["grad_p"], want["lap_u"], want["div_u"])
    got = {k: x[k].copy() for k in want}
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = {k: np.ascontiguousarray(s[k]) for k in ("imove", "r", "rho", "m")}
    emu.emu_interactions(dims, morris, P(arr["imove"]), P(arr["r"]), P(x["u"]), P(arr["rho"]), P(arr["m"]), P(x["p"]),
                         P(got["grad_p"]), P(got["lap_u"]), P(got["div_u"]), s["N"], C.c_float(D.H),
                         C.c_float(D.CONF), C.c_float(D.SUPPORT))
    fl = s["imove"] == 1
    for k in want:
        a, b = want[k].astype(np.float64), got[k].astype(np.float64)
        assert np.isfinite(b).all(), k
        assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a).max() + 2e-5 * np.abs(a)), (k, np.abs(a - b).max())
        assert np.array_equal(got[k][~fl], x[k][~fl]) and np.abs(want[k][fl] - 7.0).max() > 1e-3, k


@pytest.mark.parametrize("morris", [0, 1])
@pytest.mark.parametrize("dims,n,hfac", [(2, 60, 3.0), (3, 24, 1.3), (2, 80, 4.0)])
def test_portal_policies_match_the_oracle(oracle, emu, dims, n, hfac, morris):
    """PPortalShepard / PPortalInteractions (cfd/Boundary/Portal/Shepard.cl, Interactions.cl under both Laplacian
    definitions; written without a GPU at hand), driven by the brute-force pair loop with imirrored as the i
    filter like their launchers set it, against the oracle (bit-identical to the scripts,
    tests/test_oracle_vs_reference.py) on a state Portal/Mirror.cl::mirror prepared: the sweeps ADD to what the
    arrays hold, so the tolerance is the sweep tests' on the added part; rows that are not mirrored keep their bits."""
    import open_boundary_common as ob
    case, s, D, x = ob.portal_sweep_state(oracle, dims, n, hfac)
    want = ob.portal_sweeps_oracle(oracle, s, D, x, morris)
    got = {k: x[k].copy() for k in want}
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = {k: np.ascontiguousarray(s[k]) for k in ("imove", "rho", "m")}
    emu.emu_portal(dims, morris, P(arr["imove"]), P(x["imirrored"]), P(x["r"]), P(x["u"]), P(arr["rho"]), P(arr["m"]),
                   P(x["p"]), P(got["grad_p"]), P(got["lap_u"]), P(got["div_u"]), P(got["shepard"]), s["N"],
                   C.c_float(D.H), C.c_float(D.CONW), C.c_float(D.CONF), C.c_float(D.SUPPORT))
    for k in want:
        rows = ob.portal_rows(s, x, k)
        add_w = want[k].astype(np.float64) - x[k]
        add_g = got[k].astype(np.float64) - x[k]
        assert np.isfinite(got[k]).all() and np.array_equal(got[k][~rows], x[k][~rows]), k
        tol = 2e-6 * max(np.abs(add_w).max(), np.abs(x[k]).max()) + 2e-5 * np.abs(add_w)
        assert np.all(np.abs(add_w - add_g) <= tol), (k, np.abs(add_w - add_g).max())
        assert np.abs(add_w[rows]).max() > 1e-3, k
