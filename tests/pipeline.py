"""Test helpers: the same kernel sequence driven through the CPU oracle and
through the CUDA C-ABI, on identical inputs.  Only tests import the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

SORT1 = ["id", "iset", "imove", "r", "normal", "tangent"]          # basic/Sort.cl stage1
SORT2 = ["rho", "m", "u"]                                         # stage2 (+ rates, swapped)


def oracle_linklist_and_sort(case):
    """LinkList on r, then the permutation of basic/Sort.cl (stage1 + stage2).

    Returns the sorted state dict (numpy)."""
    dims = case["dims"]
    ll = O.linklist(case["r"], dims, case["support"], case["h"])
    inv = ll["inv_perm"]  # id_sorted
    s = dict(case)
    for k in SORT1 + SORT2:
        s[k] = O.scatter(case[k], inv)
    # Sort.cl:119-123: sorted rates land in *_in
    s["dudt_in"] = O.scatter(case["dudt"], inv)
    s["drhodt_in"] = O.scatter(case["drhodt"], inv)
    s["dudt"] = s["dudt_in"].copy()      # midpoint.xml:33-34 copies them back
    s["drhodt"] = s["drhodt_in"].copy()
    s.update(icell=ll["icell"], ihoc=ll["ihoc"], n_cells=ll["ncells"], id_sorted=inv,
             id_unsorted=ll["perm"], r_min=ll["rmin"], r_max=ll["rmax"])
    return s


def _ll(s):
    return O.make_ll(s["icell"], s["ihoc"], s["n_cells"], s["N"])


def oracle_sweeps(s, dt=1e-4):
    """One pass of the neighbour kernels of the 3-D dam-break pipeline (SURVEY
    3.2 tools 14-81) on a sorted state; returns the outputs."""
    dims, N = s["dims"], s["N"]
    V, M = O.vs(dims), O.ms(dims)
    D = O.make_defs(dims, s["h"])
    L = _ll(s)
    o = {}
    imove, r, u, rho, m = s["imove"], s["r"], s["u"].copy(), s["rho"].copy(), s["m"]
    p = np.zeros(N, np.float32)
    O.call("eos", s["iset"], imove, rho, p, s["refd"], N, s["cs"], s["p0"])
    binormal = np.zeros((N, V), np.float32)
    tangent = s["tangent"].copy()
    O.call("binormal", s["normal"], tangent, binormal, N, dims)
    o["binormal"], o["tangent"] = binormal, tangent
    nn = np.zeros(N, np.uint32)
    O.call("neighs", L, imove, nn, 100000, dims)
    o["n_neighs"] = nn
    mls = np.zeros((N, M), np.float32)
    O.call("mls", D, L, imove, r, rho, m, mls, 1)
    o["mls_raw"] = mls.copy()
    O.call("mls_inv", imove, mls, N, 1, dims)
    o["mls"] = mls
    shep = np.zeros(N, np.float32)
    O.call("shepard", D, L, 1, imove, r, rho, m, shep)
    o["shepard"] = shep
    grad_p = np.zeros((N, V), np.float32); lap_u = np.zeros((N, V), np.float32)
    div_u = np.zeros(N, np.float32)
    O.call("interactions", D, L, imove, r, u, rho, m, p, grad_p, lap_u, div_u)
    O.call("sensors", D, L, imove, r, m, u, rho, p)
    O.call("sensors_renorm", imove, shep, u, rho, p, N, dims)
    o["u_sens"], o["rho_sens"], o["p_sens"] = u.copy(), rho.copy(), p.copy()
    lpc = np.zeros((N, V), np.float32)
    O.call("dsph_full", D, L, imove, r, rho, m, p, lpc)
    o["lap_p_corr_raw"] = lpc.copy()
    lap_p = np.zeros(N, np.float32)
    O.call("dsph_lapp", D, L, imove, r, rho, m, p, lap_p)
    o["lap_p_raw"] = lap_p.copy()
    gw = np.zeros((N, V), np.float32); dub = np.zeros(N, np.float32)
    O.call("bie_interactions", D, L, imove, r, s["normal"], u, m, gw, dub)
    o["grad_w_bi"], o["div_u_bi"] = gw, dub
    o["grad_p_fluid"], o["lap_u"], o["div_u_fluid"] = grad_p.copy(), lap_u.copy(), div_u.copy()
    O.call("bie_rates", imove, rho, p, u, gw, dub, grad_p, div_u, N, dims)
    O.call("bie_p_boundary", D, L, imove, r, m, rho, p)
    o["p"] = p.copy()
    O.call("dsph_full_mls", imove, mls, lpc, N, dims)
    O.call("dsph_lapp_corr", D, L, imove, r, rho, m, lpc, lap_p)
    o["lap_p_corr"], o["lap_p"] = lpc, lap_p
    dudt = np.zeros((N, V), np.float32); drhodt = np.zeros(N, np.float32)
    O.call("rates", s["iset"], imove, rho, grad_p, lap_u, div_u, dudt, drhodt, s["visc_dyn"],
           N, s["g"], dims)
    O.call("dsph_apply", s["iset"], imove, rho, lap_p, drhodt, s["refd"], s["delta"], N, dt)
    o["grad_p"], o["div_u"] = grad_p, div_u
    o["dudt_pre"], o["drhodt"] = dudt.copy(), drhodt
    fp = np.zeros((N, V), np.float32); mp = np.zeros((N, 4), np.float32)
    O.call("bie_force_press", imove, r, s["normal"], m, p, fp, mp, s["g"] * 0, N, dims)
    o["force_p"], o["moment_p"] = fp, mp
    O.call("bie_elastic_bounce", L, imove, r, s["normal"], m, s["u"], dudt, float(dt) * 200, 0.5, dims)
    o["dudt"] = dudt
    fe = np.zeros((N, V), np.float32)
    O.call("bie_force_bound", imove, m, o["dudt_pre"], dudt, fe, N, dims)
    o["force_elastic"] = fe
    res = np.zeros(N, np.float32)
    O.call("mp_residuals", imove, m, u, s["dudt_in"], dudt, rho, p, s["drhodt_in"], drhodt, res,
           N, dims)
    o["residual"] = res
    r2 = r.copy()
    O.call("bie_pst", L, imove, r2, s["normal"], m, rho, float(D.dims), 0.5, dims)
    o["r_pst"] = r2
    dtv = np.zeros(N, np.float32)
    O.call("timestep", imove, s["u"], dtv, N, 1.0, s["dt_min"], s["courant"], s["dt_Ma"], s["h"], dims)
    o["dt_var"] = dtv
    o["dt"] = np.float32(O.lib().aqo_reduce_min(O._arg(dtv), N))
    return o


class CudaState:
    """Device mirror of a sorted state + scratch outputs, bound by variable name."""

    def __init__(self, ctx, s):
        self.ctx = ctx
        self.s = s
        self.v = {}
        dims, N = s["dims"], s["N"]
        V, M = (4 if dims == 3 else 2), (16 if dims == 3 else 4)
        for k in ("id", "iset", "imove", "r", "normal", "tangent", "rho", "m", "u", "dudt",
                  "drhodt", "dudt_in", "drhodt_in", "icell", "ihoc", "refd", "visc_dyn", "delta"):
            self.v[k] = ctx.array(s[k])
        for k in ("binormal", "grad_p", "lap_u", "lap_p_corr", "grad_w_bi", "force_p",
                  "force_elastic", "dudt_preelastic", "u_in", "r_in"):
            self.v[k] = ctx.zeros((N, V), np.float32)
        for k in ("p", "div_u", "shepard", "lap_p", "div_u_bi", "residual_midpoint", "dt_var"):
            self.v[k] = ctx.zeros(N, np.float32)
        self.v["n_neighs"] = ctx.zeros(N, np.uint32)
        self.v["mls"] = ctx.zeros((N, M), np.float32)
        self.v["moment_p"] = ctx.zeros((N, 4), np.float32)
        for k in ("N", "cs", "p0", "g", "courant", "dt_Ma", "dt_min", "h"):
            self.v[k] = s[k]
        self.v["n_cells"] = s["n_cells"]

    def run(self, script, entry="entry", **over):
        vv = dict(self.v)
        vv.update(over)
        self.ctx.launch(script, entry, vv)

    def get(self, k):
        return self.v[k].get()

    def zero(self, k):
        a = self.v[k]
        self.ctx.fill(a, np.zeros(a.elem_bytes // 4, np.float32).tobytes())

    def copy(self, dst, src):
        self.ctx.copy(self.v[dst], self.v[src])

    def set(self, k, host):
        self.v[k].set(host)

    def reduce_min(self, k):
        from aquagpusph_b200 import _lib
        return np.float32(self.ctx.reduce(_lib.OP_MIN, self.v[k]))


class RefState:
    """The same name-bound state on host arrays, run by the reference's OWN scripts
    compiled behind the shim (oracle/ref.py)."""

    def __init__(self, R, s):
        self.R = R
        self.s = s
        self.v = {}
        dims, N = s["dims"], s["N"]
        V, M = (4 if dims == 3 else 2), (16 if dims == 3 else 4)
        for k in ("id", "iset", "imove", "r", "normal", "tangent", "rho", "m", "u", "dudt",
                  "drhodt", "dudt_in", "drhodt_in", "icell", "ihoc", "refd", "visc_dyn", "delta"):
            self.v[k] = np.ascontiguousarray(s[k]).copy()
        for k in ("binormal", "grad_p", "lap_u", "lap_p_corr", "grad_w_bi", "force_p",
                  "force_elastic", "dudt_preelastic", "u_in", "r_in"):
            self.v[k] = np.zeros((N, V), np.float32)
        for k in ("p", "div_u", "shepard", "lap_p", "div_u_bi", "residual_midpoint", "dt_var"):
            self.v[k] = np.zeros(N, np.float32)
        self.v["n_neighs"] = np.zeros(N, np.uint32)
        self.v["mls"] = np.zeros((N, M), np.float32)
        self.v["moment_p"] = np.zeros((N, 4), np.float32)
        for k in ("N", "cs", "p0", "g", "courant", "dt_Ma", "dt_min", "h"):
            self.v[k] = s[k]
        self.v["n_cells"] = s["n_cells"]

    def run(self, script, entry="entry", **over):
        self.R.run(script, entry, self.s["N"], self.v, **over)

    def get(self, k):
        return self.v[k].copy()

    def zero(self, k):
        self.v[k][...] = 0

    def copy(self, dst, src):
        self.v[dst][...] = self.v[src]

    def set(self, k, host):
        self.v[k][...] = host

    def reduce_min(self, k):
        return np.float32(self.v[k].min())


def cuda_sweeps(ctx, s, dt=1e-4):
    """Same sequence as oracle_sweeps through libaquacuda (Kernel tool, by name)."""
    return named_sweeps(CudaState(ctx, s), dt)


def ref_sweeps(R, s, dt=1e-4):
    """Same sequence through the reference's own .cl scripts (oracle/_ref)."""
    return named_sweeps(RefState(R, s), dt)


def named_sweeps(c, dt=1e-4):
    s = c.s
    o = {}
    c.run("basic/EOS.cl")
    c.run("basic/Binormal.cl")
    o["binormal"], o["tangent"] = c.get("binormal"), c.get("tangent")
    c.run("basic/neighs.cl", neighs_limit=100000)
    o["n_neighs"] = c.get("n_neighs")
    c.run("basic/MLS.cl", mls_imove=1)
    o["mls_raw"] = c.get("mls")
    c.run("basic/MLS.cl", "mls_inv", mls_imove=1)
    o["mls"] = c.get("mls")
    c.run("cfd/Shepard.cl")
    o["shepard"] = c.get("shepard")
    c.run("cfd/Interactions.cl")
    c.run("cfd/Sensors.cl")
    c.run("cfd/SensorsRenormalization.cl", dt=dt)
    o["u_sens"], o["rho_sens"], o["p_sens"] = c.get("u"), c.get("rho"), c.get("p")
    c.run("cfd/deltaSPH.cl", "full")
    o["lap_p_corr_raw"] = c.get("lap_p_corr")
    c.run("cfd/deltaSPH.cl", "lapp")
    o["lap_p_raw"] = c.get("lap_p")
    c.run("cfd/Boundary/BIe/Interactions.cl")
    o["grad_w_bi"], o["div_u_bi"] = c.get("grad_w_bi"), c.get("div_u_bi")
    o["grad_p_fluid"], o["lap_u"], o["div_u_fluid"] = c.get("grad_p"), c.get("lap_u"), c.get("div_u")
    c.run("cfd/Boundary/BIe/Rates.cl")
    c.run("cfd/Boundary/BIe/Interactions.cl", "p_boundary")
    o["p"] = c.get("p")
    c.run("cfd/deltaSPH.cl", "full_mls")
    c.run("cfd/deltaSPH.cl", "lapp_corr")
    o["lap_p_corr"], o["lap_p"] = c.get("lap_p_corr"), c.get("lap_p")
    c.zero("dudt")
    c.zero("drhodt")
    c.run("cfd/Rates.cl")
    c.run("cfd/deltaSPH.cl", "deltaSPH", dt=dt)
    o["grad_p"], o["div_u"] = c.get("grad_p"), c.get("div_u")
    c.copy("dudt_preelastic", "dudt")
    o["dudt_pre"], o["drhodt"] = c.get("dudt"), c.get("drhodt")
    c.run("cfd/Boundary/BIe/Rates.cl", "force_press", forces_r=s["g"] * 0)
    o["force_p"], o["moment_p"] = c.get("force_p"), c.get("moment_p")
    # ElasticBounce reads r_in / u_in: the un-advanced state
    c.copy("r_in", "r")
    c.set("u_in", s["u"])
    c.run("cfd/Boundary/BIe/ElasticBounce.cl", dt=float(dt) * 200)
    o["dudt"] = c.get("dudt")
    c.run("cfd/Boundary/BIe/ElasticBounce.cl", "force_bound", dudt_elastic=c.v["dudt"])
    o["force_elastic"] = c.get("force_elastic")
    c.run("basic/time_scheme/midpoint.cl", "residuals")
    o["residual"] = c.get("residual_midpoint")
    c.run("cfd/Boundary/BIe/PST.cl")
    o["r_pst"] = c.get("r")
    c.set("u", s["u"])
    c.run("cfd/TimeStep.cl", dt=1.0)
    o["dt_var"] = c.get("dt_var")
    o["dt"] = c.reduce_min("dt_var")
    return o


def named_bi_sequence(c, dt=2e-4):
    """The boundary-integral (BI) part of the 2-D dam-break pipeline (BASELINE config 1,
    examples/2D/spheric_testcase5_dambreak: tools "cfd Shepard" ... "cfd elastic bounce"),
    by script name, on a CudaState or a RefState."""
    s = c.s
    o = {}
    c.run("basic/EOS.cl")
    c.run("basic/Binormal.cl")
    c.run("cfd/Boundary/BI/Shepard.cl", "compute")
    o["shepard"] = c.get("shepard")
    c.run("cfd/Interactions.cl")
    c.run("cfd/Boundary/BI/LapU.cl", "freeslip")
    o["lap_u_bi"] = c.get("lap_u")
    c.run("cfd/Boundary/BI/GradP.cl", "freeslip")
    o["grad_p_bi"] = c.get("grad_p")
    c.run("cfd/Boundary/BI/Interpolation.cl")
    o["p_interp"] = c.get("p")
    c.run("cfd/Boundary/BI/InterpolationShepard.cl")
    o["p_bound"], o["rho_bound"] = c.get("p"), c.get("rho")
    c.run("cfd/Boundary/BI/Interactions.cl")
    o["grad_p_sum"], o["div_u_sum"] = c.get("grad_p"), c.get("div_u")
    c.run("cfd/Boundary/BI/Shepard.cl", "apply")
    o["grad_p"], o["lap_u"], o["div_u"] = c.get("grad_p"), c.get("lap_u"), c.get("div_u")
    c.zero("dudt")
    c.zero("drhodt")
    c.run("cfd/Rates.cl")
    o["dudt_pre"] = c.get("dudt")
    c.run("cfd/Boundary/ElasticBounce.cl", dr=s["dr"], dt=float(dt) * 50)
    o["u_bounce"], o["dudt_bounce"] = c.get("u"), c.get("dudt")
    return o


# tolerance per output: |a - b| <= atol_rel * max|b| + rtol * |b|
EXACT = {"n_neighs", "binormal", "tangent", "dt_var", "dt"}
ORDER_DEP = {"dudt", "force_elastic", "r_pst", "residual"}


def compare(o_ref, o_gpu, rtol=2e-5, atol_rel=2e-6):
    """Returns list of (name, max_abs_err, scale, ok)."""
    rep = []
    for k, a in o_ref.items():
        b = o_gpu[k]
        a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
        scale = float(np.max(np.abs(a64))) if a64.size else 0.0
        err = np.abs(a64 - b64)
        if k in EXACT:
            ok = bool(np.array_equal(np.asarray(a), np.asarray(b)))
        else:
            ok = bool(np.all(err <= atol_rel * scale + rtol * np.abs(a64)))
        rep.append((k, float(err.max()) if err.size else 0.0, scale, ok))
    return rep
