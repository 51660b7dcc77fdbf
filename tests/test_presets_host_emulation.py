"""CPU check of the LOGIC of the CUDA kernels behind cfd/motion.xml, cfd/energy.xml and the small
presets next to them: the kernel bodies of aquagpusph_b200/csrc/elementwise.cu (k_motion_*,
k_energy_*, k_ab_*, k_forces, k_density_clamp, k_id_inverse, with the V<D> helpers they use) are lifted out of the .cu file as text, compiled for the host by g++ behind a two-screen shim
(float2/float4, __global__ = nothing, the thread index as a loop variable) without FMA contraction,
and compared with the oracle on the inputs of tests/test_gpu_presets.py.

What this pins: index and sign conventions, operation order, the host-side cos/sin hoisting.  What
it cannot pin: the device's own arithmetic (IEEE for + - * / sqrt, so identical; logf is not) and
the launchers' argument slots -- tests/test_gpu_presets.py does that on a B200.  Nothing here is a
product path: the shim exists only in this test's temporary directory."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CU = os.path.join(ROOT, "aquagpusph_b200", "csrc", "elementwise.cu")

SHIM = r"""
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
struct aqc_f4 { float x, y, z, w; };
#define __device__
#define __global__
#define __launch_bounds__(n)
static size_t g_i;
#define GID const size_t i = g_i; if (i >= N) return;
"""

WRAP = r"""
static aqc_f4 f4(const float* p, int n) { aqc_f4 v{ 0.f, 0.f, 0.f, 0.f }; memcpy(&v, p, 4 * n); return v; }
#define FOR_ALL(CALL3, CALL2) for (g_i = 0; g_i < N; g_i++) { if (dims == 3) { CALL3; } else { CALL2; } }
extern "C" {
void emu_transform(int dims, const uint32_t* iset, const int* imove, void* r, void* n, void* t, uint32_t N,
                   uint32_t set, const float* mr, const float* ang)
{
    aqc_f4 a = f4(ang, 4), lin = f4(mr, dims == 3 ? 4 : 2);
    FOR_ALL(k_motion_transform<3>(iset, imove, r, n, t, N, set, lin, motion_trig(a, 1.f)),
            k_motion_transform<2>(iset, imove, r, n, t, N, set, lin, motion_trig(a, 1.f)))
}
void emu_untransform(int dims, const uint32_t* iset, const int* imove, void* r, void* n, void* t, uint32_t N,
                     uint32_t set, const float* mr, const float* ang)
{
    aqc_f4 a = f4(ang, 4), lin = f4(mr, dims == 3 ? 4 : 2);
    FOR_ALL(k_motion_untransform<3>(iset, imove, r, n, t, N, set, lin, motion_trig(a, -1.f)),
            k_motion_untransform<2>(iset, imove, r, n, t, N, set, lin, motion_trig(a, -1.f)))
}
void emu_rate(int dims, const uint32_t* iset, const int* imove, const void* r, void* out, uint32_t N,
              uint32_t set, const float* lin_, const float* ang, const float* w_)
{
    aqc_f4 a = f4(ang, 4), w = f4(w_, 4), lin = f4(lin_, dims == 3 ? 4 : 2);
    FOR_ALL(k_motion_rate<3>(iset, imove, r, out, N, set, lin, w, motion_trig(a, 1.f)),
            k_motion_rate<2>(iset, imove, r, out, N, set, lin, w, motion_trig(a, 1.f)))
}
void emu_power(int dims, float* dek, float* dep, float* dec, const int* imove, const void* u, const float* rho,
               const float* m, const float* p, const void* dudt, const float* drhodt, uint32_t N, const float* g_)
{
    aqc_f4 g = f4(g_, dims == 3 ? 4 : 2);
    FOR_ALL(k_energy_power<3>(dek, dep, dec, imove, u, rho, m, p, dudt, drhodt, N, g),
            k_energy_power<2>(dek, dep, dec, imove, u, rho, m, p, dudt, drhodt, N, g))
}
void emu_energy(int dims, float* ek, float* ep, float* ec, const uint32_t* iset, const int* imove, const void* r,
                const void* u, const float* rho, const float* m, const float* refd, uint32_t N, const float* g_,
                float cs)
{
    aqc_f4 g = f4(g_, dims == 3 ? 4 : 2);
    FOR_ALL(k_energy_energy<3>(ek, ep, ec, iset, imove, r, u, rho, m, refd, N, g, cs),
            k_energy_energy<2>(ek, ep, ec, iset, imove, r, u, rho, m, refd, N, g, cs))
}
void emu_ab_step(int dims, void* const* du, void* const* du_in, float* const* dr, float* const* dr_in,
                 const uint32_t* id_sorted, const int* imove, void* r, void* u, const void* dudt, float* rho,
                 const float* drhodt, uint32_t N, float dt, unsigned iter, unsigned steps)
{
    AB4 lv, in;
    AB4c lvc, inc;
    for (int l = 0; l < 4; l++) {
        lv.du[l] = du[l]; lv.dr[l] = dr[l]; lvc.du[l] = du[l]; lvc.dr[l] = dr[l];
        in.du[l] = du_in[l]; in.dr[l] = dr_in[l]; inc.du[l] = du_in[l]; inc.dr[l] = dr_in[l];
    }
    const unsigned local_iter = iter < steps ? iter : steps;   // l_ab_corrector
    FOR_ALL(k_ab_sort<3>(inc, lv, id_sorted, N), k_ab_sort<2>(inc, lv, id_sorted, N))
    FOR_ALL(k_ab_corrector<3>(imove, r, u, dudt, rho, drhodt, lvc, N, dt, local_iter),
            k_ab_corrector<2>(imove, r, u, dudt, rho, drhodt, lvc, N, dt, local_iter))
    FOR_ALL(k_ab_postcorrector<3>(lvc, dudt, drhodt, in, N), k_ab_postcorrector<2>(lvc, dudt, drhodt, in, N))
}
void emu_small(int dims, float* ekin, void* ff, float4* fm, float* rho_in, const uint32_t* id, uint32_t* inv,
               const int* imove, const void* r, const void* u, const void* dudt, const float* m, uint32_t N,
               const float* g_, const float* fr_, float lo, float hi)
{
    aqc_f4 g = f4(g_, dims == 3 ? 4 : 2), fr = f4(fr_, dims == 3 ? 4 : 2);
    FOR_ALL(k_energy_kin<3>(ekin, imove, u, m, N), k_energy_kin<2>(ekin, imove, u, m, N))
    FOR_ALL(k_forces<3>(ff, fm, imove, r, dudt, m, N, g, fr), k_forces<2>(ff, fm, imove, r, dudt, m, N, g, fr))
    for (g_i = 0; g_i < N; g_i++) {
        k_density_clamp(rho_in, N, lo, hi);
        k_id_inverse(id, inv, N);
    }
}
/* cfd/Boundary/Symmetry/Mirror.cl: detect, (the preset's radix sort happens outside), feed, set, sort, drop */
void emu_open_boundary(int dims, int step, int* imove, const uint32_t* iset, void* r, void* r_in, void* u, void* dudt,
                       void* dudt_in, float* rho, float* drhodt, float* drhodt_in, float* m, float* p,
                       const float* refd, int* imirrored, uint32_t* icell, uint32_t N_, uint32_t nbuffer,
                       const float* vecs, const float* fl, const uint32_t* un, float support_h)
{
    // vecs: g, domain_max, inlet_r, inlet_ru, inlet_rv, inlet_n, inlet_rFS, outlet_r, outlet_n, outlet_rFS,
    //       portal_in_r, portal_out_r, portal_n, r_min (4 floats each)
    // fl: cs, p0, dr, inlet_U, inlet_R, outlet_U;  un: inlet_N.x, inlet_N.y, n_cells.x, n_cells.y
    const int w = dims == 3 ? 4 : 2;
#define VV(k) f4(vecs + 4 * (k), w)
    uint32_t N = N_;
    if (step == 0) { // the launcher's arithmetic (l_inlet_feed)
        const uint64_t want = (uint64_t)un[0] * un[1];
        const uint32_t count = (uint32_t)(want < nbuffer ? want : nbuffer);
        const float off = fl[4] - support_h - 0.5f * fl[2];
        const uint32_t first = N_ - nbuffer;
        N = count;
        FOR_ALL(k_inlet_feed<3>(imove, iset, r, u, dudt, rho, drhodt, m, p, refd, first, N, fl[0], fl[1], VV(0), fl[2],
                                VV(2), VV(3), VV(4), un[0], un[1], VV(5), fl[3], VV(6), off),
                k_inlet_feed<2>(imove, iset, r, u, dudt, rho, drhodt, m, p, refd, first, N, fl[0], fl[1], VV(0), fl[2],
                                VV(2), VV(3), VV(4), un[0], un[1], VV(5), fl[3], VV(6), off))
    } else if (step == 1) {
        FOR_ALL(k_inlet_rates<3>(imove, r, u, dudt, drhodt, N, VV(2), fl[3], VV(5)),
                k_inlet_rates<2>(imove, r, u, dudt, drhodt, N, VV(2), fl[3], VV(5)))
    } else if (step == 2) {
        FOR_ALL(k_outlet_rates<3>(imove, iset, r, u, rho, p, dudt, dudt_in, drhodt, drhodt_in, refd, N, fl[0], fl[1],
                                  VV(0), VV(7), VV(8), fl[5], VV(9)),
                k_outlet_rates<2>(imove, iset, r, u, rho, p, dudt, dudt_in, drhodt, drhodt_in, refd, N, fl[0], fl[1],
                                  VV(0), VV(7), VV(8), fl[5], VV(9)))
    } else if (step == 3) {
        FOR_ALL(k_outlet_feed<3>(imove, r_in, N, VV(1), VV(7), VV(8), support_h),
                k_outlet_feed<2>(imove, r_in, N, VV(1), VV(7), VV(8), support_h))
    } else if (step == 4) {
        FOR_ALL(k_portal_mirror<3>(r, imirrored, icell, N, VV(10), VV(11), VV(12), VV(13), un[2], un[3], support_h,
                                   1.f / support_h),
                k_portal_mirror<2>(r, imirrored, icell, N, VV(10), VV(11), VV(12), VV(13), un[2], un[3], support_h,
                                   1.f / support_h))
    } else if (step == 5) {
        FOR_ALL(k_portal_unmirror<3>(r, imirrored, N, VV(10), VV(11)), k_portal_unmirror<2>(r, imirrored, N, VV(10), VV(11)))
    } else {
        FOR_ALL(k_portal_teleport<3>(r, N, VV(10), VV(11), VV(12)), k_portal_teleport<2>(r, N, VV(10), VV(11), VV(12)))
    }
#undef VV
}
void emu_sym_detect(int dims, const int* imove, const void* r_in, uint32_t* imirror, uint32_t N, const float* sr_,
                    const float* sn_, float support_h)
{
    aqc_f4 sr = f4(sr_, dims == 3 ? 4 : 2), sn = f4(sn_, dims == 3 ? 4 : 2);
    FOR_ALL(k_sym_detect<3>(imove, r_in, imirror, N, sr, sn, support_h),
            k_sym_detect<2>(imove, r_in, imirror, N, sr, sn, support_h))
}
void emu_sym_rest(int dims, int* imove, uint32_t* iset, const uint32_t* imirror, const uint32_t* invperm,
                  uint32_t* mirror_src, uint32_t* mirror_src_in, const uint32_t* id_sorted, void* normal, void* tangent,
                  void* r_in, void* r, float* m, void* u_in, void* dudt_in, void* dudt, float* rho_in, float* drhodt_in,
                  float* drhodt, uint32_t N, uint32_t nbuffer, const float* sr_, const float* sn_, const float* dmax_)
{
    const int n = dims == 3 ? 4 : 2;
    aqc_f4 sr = f4(sr_, n), sn = f4(sn_, n), dmax = f4(dmax_, n);
    FOR_ALL(k_sym_feed<3>(imove, iset, imirror, invperm, mirror_src, normal, tangent, r_in, N, nbuffer, sr, sn),
            k_sym_feed<2>(imove, iset, imirror, invperm, mirror_src, normal, tangent, r_in, N, nbuffer, sr, sn))
    FOR_ALL(k_sym_set<3>(mirror_src, m, u_in, dudt_in, dudt, rho_in, drhodt_in, drhodt, N, sn),
            k_sym_set<2>(mirror_src, m, u_in, dudt_in, dudt, rho_in, drhodt_in, drhodt, N, sn))
    memcpy(mirror_src_in, mirror_src, 4 * (size_t)N);
    for (g_i = 0; g_i < N; g_i++)
        k_sym_sort(mirror_src_in, mirror_src, id_sorted, N);
    memcpy(r, r_in, 4 * (size_t)n * N);
    FOR_ALL(k_sym_drop<3>(imove, r, N, sr, sn, dmax), k_sym_drop<2>(imove, r, N, sr, sn, dmax))
}
}
"""


def _lift():
    """The V<D> helpers and the motion / energy kernels of elementwise.cu, launchers removed."""
    src = open(CU).read()
    a = src.index("template <int D> struct V;")
    b = src.index("#define GID")
    helpers = src[a:b]
    a = src.index("// ---- cfd/Motions/")
    b = src.index("// ---- basic/Sort.cl")
    body = src[a:b]
    # launchers: 'int l_xxx(aqc_ctx* c, ...)\n{ ... \n}\n' at column 0
    body = re.sub(r"^int l_\w+\(aqc_ctx\*[^\n]*\n\{\n.*?^\}\n", "", body, flags=re.S | re.M)
    assert "DISPATCH" not in body and "LAUNCH(" not in body
    assert all(k in body for k in ("k_energy_energy", "k_motion_rate", "k_forces", "k_id_inverse", "k_ab_corrector",
                                   "k_inlet_feed", "k_outlet_rates", "k_portal_mirror"))
    return helpers + body


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu")
    cpp, so = str(d / "emu.cpp"), str(d / "libemu.so")
    open(cpp, "w").write(SHIM + _lift() + WRAP)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                           "-fno-fast-math", "-o", so, cpp])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("dims", [2, 3])
def test_motion_kernel_bodies_match_the_oracle(oracle, emu, dims):
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
    mr = np.array([0.3, -0.2, 0.1, 0.0], np.float32)[:V].copy()
    ma = np.array([0.21, -0.13, 0.37, 0.0], np.float32)
    drdt = np.array([0.5, 0.25, -0.125, 0.0], np.float32)[:V].copy()
    dadt = np.array([0.7, -0.4, 1.1, 0.0], np.float32)
    ddr = np.array([-1.5, 0.75, 2.0, 0.0], np.float32)[:V].copy()
    dda = np.array([0.9, 0.3, -0.6, 0.0], np.float32)
    e = {k: v.copy() for k, v in h.items()}
    o = {k: v.copy() for k, v in h.items()}

    def same(keys):
        for k in keys:
            assert e[k].tobytes() == o[k].tobytes(), k

    emu.emu_rate(dims, _p(e["iset"]), _p(e["imove"]), _p(e["r"]), _p(e["u"]), N, 1, _p(drdt), _p(ma), _p(dadt))
    oracle.call("motion_velocity", o["iset"], o["imove"], o["r"], o["u"], N, 1, drdt, ma, dadt, dims)
    emu.emu_rate(dims, _p(e["iset"]), _p(e["imove"]), _p(e["r"]), _p(e["dudt"]), N, 1, _p(ddr), _p(ma), _p(dda))
    oracle.call("motion_acceleration", o["iset"], o["imove"], o["r"], o["dudt"], N, 1, ddr, ma, dda, dims)
    same(("u", "dudt"))
    assert np.abs(o["u"]).max() > 0.1 and np.abs(o["dudt"]).max() > 0.1
    emu.emu_transform(dims, _p(e["iset"]), _p(e["imove"]), _p(e["r"]), _p(e["normal"]), _p(e["tangent"]), N, 1,
                      _p(mr), _p(ma))
    oracle.call("motion_transform", o["iset"], o["imove"], o["r"], o["normal"], o["tangent"], N, 1, mr, ma, dims)
    same(("r", "normal", "tangent"))
    moved = (h["iset"] == 1) & (h["imove"] != 1)
    assert moved.any() and np.abs(o["r"][moved] - h["r"][moved]).max() > 0.05
    emu.emu_untransform(dims, _p(e["iset"]), _p(e["imove"]), _p(e["r"]), _p(e["normal"]), _p(e["tangent"]), N, 1,
                        _p(mr), _p(ma))
    oracle.call("motion_untransform", o["iset"], o["imove"], o["r"], o["normal"], o["tangent"], N, 1, mr, ma, dims)
    same(("r", "normal", "tangent"))


@pytest.mark.parametrize("dims", [2, 3])
def test_energy_kernel_bodies_match_the_oracle(oracle, emu, dims):
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
    cs = float(case["cs"])
    names = ("energy_dekdt", "energy_depdt", "energy_decdt", "energy_ek", "energy_ep", "energy_ec")
    o = {k: np.full(N, 7.0, np.float32) for k in names}
    e = {k: np.full(N, 7.0, np.float32) for k in names}
    oracle.call("energy_power", o["energy_dekdt"], o["energy_depdt"], o["energy_decdt"], v["imove"], v["u"],
                v["rho"], v["m"], v["p"], v["dudt"], v["drhodt"], N, g, dims)
    oracle.call("energy_energy", o["energy_ek"], o["energy_ep"], o["energy_ec"], v["iset"], v["imove"], v["r"],
                v["u"], v["rho"], v["m"], v["refd"], N, g, cs, dims)
    emu.emu_power(dims, _p(e["energy_dekdt"]), _p(e["energy_depdt"]), _p(e["energy_decdt"]), _p(v["imove"]),
                  _p(v["u"]), _p(v["rho"]), _p(v["m"]), _p(v["p"]), _p(v["dudt"]), _p(v["drhodt"]), N, _p(g))
    emu.emu_energy(dims, _p(e["energy_ek"]), _p(e["energy_ep"]), _p(e["energy_ec"]), _p(v["iset"]),
                   _p(v["imove"]), _p(v["r"]), _p(v["u"]), _p(v["rho"]), _p(v["m"]), _p(v["refd"]), N, _p(g),
                   C.c_float(cs))
    for k in names:
        assert e[k].tobytes() == o[k].tobytes(), k   # the same libm logf on both sides here
        assert np.abs(o[k]).max() > 0


@pytest.mark.parametrize("dims", [2, 3])
def test_small_preset_kernel_bodies_match_the_oracle(oracle, emu, dims):
    """k_energy_kin, k_forces, k_density_clamp, k_id_inverse (cfd/energy_kin.xml, cfd/forces.xml,
    basic/densityClamp.xml, basic/id_inverse.xml)."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(8)
    v = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "r", "m", "rho")}
    v["u"] = rng.normal(size=(N, V)).astype(np.float32)
    v["dudt"] = rng.normal(size=(N, V)).astype(np.float32)
    g = np.asarray(case["g"], np.float32).ravel()[:V].copy()
    fr = np.array([0.3, -0.1, 0.2, 0.0], np.float32)[:V].copy()
    perm = rng.permutation(N).astype(np.uint32)
    lo, hi = float(np.percentile(v["rho"], 20)), float(np.percentile(v["rho"], 80))
    o = dict(rho_in=v["rho"].copy(), energy_kin=np.full(N, 7.0, np.float32),
             forces_f=np.full((N, V), 7.0, np.float32), forces_m=np.full((N, 4), 7.0, np.float32),
             id_inverse=np.zeros(N, np.uint32))
    e = {k: a.copy() for k, a in o.items()}
    oracle.call("energy_kin", o["energy_kin"], v["imove"], v["u"], v["m"], N, dims)
    oracle.call("forces", o["forces_f"], o["forces_m"], v["imove"], v["r"], v["dudt"], v["m"], N, g, fr, dims)
    oracle.call("density_clamp", o["rho_in"], N, lo, hi)
    oracle.call("id_inverse", perm, o["id_inverse"], N)
    emu.emu_small(dims, _p(e["energy_kin"]), _p(e["forces_f"]), _p(e["forces_m"]), _p(e["rho_in"]), _p(perm),
                  _p(e["id_inverse"]), _p(v["imove"]), _p(v["r"]), _p(v["u"]), _p(v["dudt"]), _p(v["m"]), N,
                  _p(g), _p(fr), C.c_float(lo), C.c_float(hi))
    for k in o:
        assert e[k].tobytes() == o[k].tobytes(), k
    assert np.abs(o["forces_m"][:, 2]).max() > 0 and np.abs(o["energy_kin"]).max() > 0


@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("steps", [5, 2])
def test_adams_bashforth_kernel_bodies_match_the_oracle(oracle, emu, dims, steps):
    """k_ab_sort / k_ab_corrector / k_ab_postcorrector over seven steps (every order), also with
    TSCHEME_ADAMS_BASHFORTH_STEPS = 2 capping the order."""
    from test_oracle_vs_reference import _ab_state, ab_oracle_step
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N = case["N"]
    o = _ab_state(case, dims, 21)
    e = {k: x.copy() for k, x in o.items()}
    PP = lambda keys: (C.c_void_p * 4)(*[e[k].ctypes.data for k in keys])   # noqa: E731
    lv = range(1, 5)
    for it in range(7):
        ab_oracle_step(oracle, o, N, dims, 1.25e-3, it, steps)
        emu.emu_ab_step(dims, PP("dudt_as%d" % l for l in lv), PP("dudt_as%d_in" % l for l in lv),
                        PP("drhodt_as%d" % l for l in lv), PP("drhodt_as%d_in" % l for l in lv),
                        _p(e["id_sorted"]), _p(e["imove"]), _p(e["r"]), _p(e["u"]), _p(e["dudt"]), _p(e["rho"]),
                        _p(e["drhodt"]), N, C.c_float(1.25e-3), it, steps)
        for k in o:
            assert e[k].tobytes() == o[k].tobytes(), (it, k)
        o["dudt"][:, :dims] = np.random.default_rng(100 + it).normal(size=(N, dims)).astype(np.float32)
        e["dudt"][...] = o["dudt"]


@pytest.mark.parametrize("dims", [2, 3])
def test_symmetry_mirror_kernel_bodies_match_the_oracle(oracle, emu, dims):
    """k_sym_detect / feed / set / sort / drop (cfd/Boundary/Symmetry/Mirror.cl, preset cfd/symmetry.xml) in
    the preset's order, on the state of tests/test_oracle_vs_reference.py's symmetry test."""
    import test_oracle_vs_reference as T
    case, v, N, nbuf, sr, sn, dmax = T._symmetry_state(dims)
    V = 4 if dims == 3 else 2
    D = oracle.make_defs(dims, case["h"])

    def state():
        return dict(imove=v["imove"].copy(), iset=v["iset"].astype(np.uint32), r_in=v["r"].copy(), r=v["r"].copy(),
                    normal=v["normal"].copy(), tangent=v["tangent"].copy(), m=v["m"].copy(), u_in=v["u"].copy(),
                    dudt_in=v["dudt"].copy(), dudt=np.zeros((N, V), np.float32), rho_in=v["rho"].copy(),
                    drhodt_in=v["drhodt"].copy(), drhodt=np.zeros(N, np.float32),
                    imirror=np.full(N, 7, np.uint32), mirror_src=np.full(N, N, np.uint32),
                    mirror_src_in=np.zeros(N, np.uint32))
    e, o = state(), state()
    emu.emu_sym_detect(dims, _p(e["imove"]), _p(e["r_in"]), _p(e["imirror"]), N, _p(sr), _p(sn),
                       C.c_float(float(np.float32(D.SUPPORT) * np.float32(D.H))))
    oracle.call("sym_detect", D, o["imove"], o["r_in"], o["imirror"], N, sr, sn)
    assert np.array_equal(e["imirror"], o["imirror"]) and e["imirror"].sum() > 0
    perm = np.argsort(o["imirror"], kind="stable").astype(np.uint32)
    inv = np.empty(N, np.uint32)
    inv[perm] = np.arange(N, dtype=np.uint32)
    ids = np.random.default_rng(4).permutation(N).astype(np.uint32)
    for d in (e, o):
        d["imirror"] = d["imirror"][perm].copy()
    emu.emu_sym_rest(dims, _p(e["imove"]), _p(e["iset"]), _p(e["imirror"]), _p(inv), _p(e["mirror_src"]),
                     _p(e["mirror_src_in"]), _p(ids), _p(e["normal"]), _p(e["tangent"]), _p(e["r_in"]), _p(e["r"]),
                     _p(e["m"]), _p(e["u_in"]), _p(e["dudt_in"]), _p(e["dudt"]), _p(e["rho_in"]), _p(e["drhodt_in"]),
                     _p(e["drhodt"]), N, nbuf, _p(sr), _p(sn), _p(dmax))
    oi = o["iset"].view(np.int32)
    oracle.call("sym_feed", o["imove"], oi, o["imirror"], inv, o["mirror_src"], o["normal"], o["tangent"], o["r_in"],
                N, nbuf, sr, sn, dims)
    oracle.call("sym_set", o["mirror_src"], o["m"], o["u_in"], o["dudt_in"], o["dudt"], o["rho_in"], o["drhodt_in"],
                o["drhodt"], N, sn, dims)
    o["mirror_src_in"] = o["mirror_src"].copy()
    oracle.call("sym_sort", o["mirror_src_in"], o["mirror_src"], ids, N)
    o["r"] = o["r_in"].copy()
    oracle.call("sym_drop", o["imove"], o["r"], N, sr, sn, dmax, dims)
    for k in e:
        assert e[k].tobytes() == o[k].tobytes(), k
    assert (o["mirror_src_in"] < N).sum() == int(o["imirror"].sum()) > 0 and (o["imove"] == -256).sum() > 0


@pytest.mark.parametrize("dims", [2, 3])
def test_open_boundary_kernel_bodies_match_the_oracle(oracle, emu, dims):
    """k_inlet_* / k_outlet_* / k_portal_* of elementwise.cu (cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl,
    Portal/Mirror.cl) with their launchers' arithmetic, on the state and in the order of
    tests/test_oracle_vs_reference.py::test_open_boundary_kernels_match_reference_scripts, which pins the
    oracle to the reference's scripts: the same bits after every kernel."""
    import open_boundary_common as ob
    case, v = ob.state(dims)
    D = oracle.make_defs(dims, case["h"])
    e, o = ob.args_of(v), ob.args_of(v)

    def pad(x):
        a = np.zeros(4, np.float32)
        a[:len(x)] = x
        return a
    vecs = np.concatenate([pad(v[k]) for k in ("g", "domain_max", "inlet_r", "inlet_ru", "inlet_rv", "inlet_n",
                                               "inlet_rFS", "outlet_r", "outlet_n", "outlet_rFS", "portal_in_r",
                                               "portal_out_r", "portal_n", "r_min")])
    fl = np.array([v["cs"], v["p0"], v["dr"], v["inlet_U"], v["inlet_R"], v["outlet_U"]], np.float32)
    un = np.array([v["inlet_N"][0], v["inlet_N"][1], v["n_cells"][0], v["n_cells"][1]], np.uint32)
    for step, key in enumerate(ob.STEPS):
        emu.emu_open_boundary(dims, step, _p(e["imove"]), _p(e["iset"]), _p(e["r"]), _p(e["r_in"]), _p(e["u"]),
                              _p(e["dudt"]), _p(e["dudt_in"]), _p(e["rho"]), _p(e["drhodt"]), _p(e["drhodt_in"]),
                              _p(e["m"]), _p(e["p"]), _p(e["refd"]), _p(e["imirrored"]), _p(e["icell"]),
                              C.c_uint32(v["N"]), C.c_uint32(v["nbuffer"]), _p(vecs), _p(fl), _p(un),
                              C.c_float(D.SUPPORT * D.H))
        ob.oracle_step(oracle, D, dims, key, o)
        for k in ob.WRITES[key]:
            assert e[k].tobytes() == o[k].tobytes(), (key, k)
    for k in e:
        if isinstance(e[k], np.ndarray):
            assert e[k].tobytes() == o[k].tobytes(), k
    ob.checks(v, o, dims)


# ---- cfd/ideal_gas: the element-wise kernels ---------------------------------------------------------------
IG_WRAP = r"""
extern "C" void emu_ig(int dims, const uint32_t* iset, const int* imove, const float* rho, float* eint, float* p,
                       const float* gamma, const float* div_u, float* deintdt, float* dt_var, const void* u,
                       const void* grad_p, const float* work_density, float* eint_in, float* deintdt_in,
                       const uint32_t* id_sorted, const uint32_t* mirror_src, uint32_t N, float dt, float dt_min,
                       float courant, float H, float relax)
{
    aqc_sv<float> rx{ nullptr, relax };
    for (g_i = 0; g_i < N; g_i++) k_ig_eos(iset, imove, rho, eint, p, gamma, N);
    for (g_i = 0; g_i < N; g_i++) k_ig_rates(imove, rho, p, div_u, deintdt, N);
    FOR_ALL(k_ig_timestep<3>(dt_var, imove, iset, u, rho, p, N, dt, dt_min, courant, div_u, grad_p, gamma, H),
            k_ig_timestep<2>(dt_var, imove, iset, u, rho, p, N, dt, dt_min, courant, div_u, grad_p, gamma, H))
    for (g_i = 0; g_i < N; g_i++) k_ig_mp_predictor(eint, deintdt, eint_in, deintdt_in, N);
    for (g_i = 0; g_i < N; g_i++) k_ig_riemann_rates(imove, work_density, deintdt, N);
    for (g_i = 0; g_i < N; g_i++) k_ig_mp_advance(imove, eint_in, deintdt, eint, N, 0.5f * dt);
    for (g_i = 0; g_i < N; g_i++) k_ig_mp_relax(imove, deintdt_in, deintdt, N, rx);
    for (g_i = 0; g_i < N; g_i++) k_ig_mp_advance(imove, eint_in, deintdt, eint, N, dt);
    memcpy(eint_in, eint, 4 * (size_t)N);
    for (g_i = 0; g_i < N; g_i++) k_ig_sort(eint_in, eint, deintdt, deintdt_in, id_sorted, N);
    for (g_i = 0; g_i < N; g_i++) k_ig_mp_predictor(eint, deintdt, eint_in, deintdt_in, N);
    for (g_i = 0; g_i < N; g_i++) k_ig_euler_corrector(imove, eint, deintdt, N, dt);
    memcpy(deintdt, work_density, 4 * (size_t)N);
    for (g_i = 0; g_i < N; g_i++) k_ig_ie_corrector(imove, deintdt, deintdt_in, eint, N, dt);
    for (g_i = 0; g_i < N; g_i++) k_ig_ie_predictor(imove, eint, deintdt, eint_in, deintdt_in, N, dt);
    for (g_i = 0; g_i < N; g_i++) k_ig_sym_set(mirror_src, eint_in, deintdt_in, deintdt, N);
}
"""


@pytest.fixture(scope="module")
def emu_ig(tmp_path_factory):
    src = open(CU).read()
    a = src.index("template <int D> struct V;")
    helpers = src[a:src.index("#define GID")]
    a = src.index("// ---- cfd/ideal_gas:")
    body = src[a:src.index("#define IN(n, t)")]
    body = re.sub(r"^int l_\w+\(aqc_ctx\*[^\n]*\n\{\n.*?^\}\n", "", body, flags=re.S | re.M)
    assert "LAUNCH(" not in body and "DISPATCH" not in body and "k_ig_mp_relax" in body
    sv = "template <typename T> struct aqc_sv { const T* p; T v; T get() const { return p ? *p : v; } };\n"
    d = tmp_path_factory.mktemp("emu_ig")
    cpp, so = str(d / "emu.cpp"), str(d / "libemu.so")
    open(cpp, "w").write(SHIM + sv + helpers + body +
                         "#define FOR_ALL(CALL3, CALL2) for (g_i = 0; g_i < N; g_i++) { if (dims == 3) { CALL3; } "
                         "else { CALL2; } }\n" + IG_WRAP)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                           "-fno-fast-math", "-o", so, cpp])
    return C.CDLL(so)


@pytest.mark.parametrize("dims", [2, 3])
def test_ideal_gas_kernel_bodies_match_the_oracle(oracle, emu_ig, dims):
    """k_ig_* of elementwise.cu (cfd/ideal_gas/{EOS, Rates, Sort, TimeStep}.cl, riemann/Rates.cl,
    time_scheme/midpoint.cl) in the order of tests/test_oracle_vs_reference.py, which pins the oracle to
    the reference's scripts: the same bits."""
    from test_oracle_vs_reference import _ideal_gas_state
    import oracle.oracle as O
    e, o, N = _ideal_gas_state(dims, 13)
    D = O.make_defs(dims, o["h"])
    oracle.call("ig_eos", o["iset"], o["imove"], o["rho"], o["eint"], o["p"], o["gamma"], N)
    oracle.call("ig_rates", o["imove"], o["rho"], o["p"], o["div_u"], o["deintdt"], N)
    oracle.call("ig_timestep", D, o["dt_var"], o["imove"], o["iset"], o["u"], o["rho"], o["p"], N, o["dt"],
                o["dt_min"], o["courant"], o["div_u"], o["grad_p"], o["gamma"])
    oracle.call("ig_mp_predictor", o["eint"], o["deintdt"], o["eint_in"], o["deintdt_in"], N)
    oracle.call("ig_riemann_rates", o["imove"], o["work_density"], o["deintdt"], N)
    oracle.call("ig_mp_midpoint", o["imove"], o["eint_in"], o["deintdt"], o["eint"], N, o["dt"])
    oracle.call("ig_mp_relax", o["imove"], o["deintdt_in"], o["deintdt"], N, o["relax_midpoint"])
    oracle.call("ig_mp_corrector", o["imove"], o["eint_in"], o["deintdt"], o["eint"], N, o["dt"])
    o["eint_in"][...] = o["eint"]
    oracle.call("ig_sort", o["eint_in"], o["eint"], o["deintdt"], o["deintdt_in"], o["id_sorted"], N)
    oracle.call("ig_mp_predictor", o["eint"], o["deintdt"], o["eint_in"], o["deintdt_in"], N)
    oracle.call("ig_euler_corrector", o["imove"], o["eint"], o["deintdt"], N, o["dt"])
    o["deintdt"][...] = o["work_density"]
    oracle.call("ig_ie_corrector", o["imove"], o["deintdt"], o["deintdt_in"], o["eint"], N, o["dt"])
    oracle.call("ig_ie_predictor", o["imove"], o["eint"], o["deintdt"], o["eint_in"], o["deintdt_in"], N, o["dt"])
    oracle.call("ig_sym_set", o["mirror_src"], o["eint_in"], o["deintdt_in"], o["deintdt"], N)
    emu_ig.emu_ig(dims, _p(e["iset"]), _p(e["imove"]), _p(e["rho"]), _p(e["eint"]), _p(e["p"]), _p(e["gamma"]),
                  _p(e["div_u"]), _p(e["deintdt"]), _p(e["dt_var"]), _p(e["u"]), _p(e["grad_p"]),
                  _p(e["work_density"]), _p(e["eint_in"]), _p(e["deintdt_in"]), _p(e["id_sorted"]), _p(e["mirror_src"]), N,
                  C.c_float(e["dt"]), C.c_float(e["dt_min"]), C.c_float(e["courant"]), C.c_float(D.H),
                  C.c_float(e["relax_midpoint"]))
    for k in ("p", "deintdt", "dt_var", "eint", "eint_in", "deintdt_in"):
        assert e[k].tobytes() == o[k].tobytes(), k
