"""Shared inputs of the open-boundary kernel tests (cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl,
Portal/Mirror.cl; presets cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml): one state per dimension, the
kernel sequence as (script, entry, argument dict) steps, and the oracle run of it.  The CPU tests
(reference scripts, committed outputs, host emulation of the CUDA bodies) and the GPU test run the same
steps on the same state."""
import numpy as np

import cases

INLET, OUTLET, PORTAL = ("cfd/Boundary/Inlet/Inlet.cl", "cfd/Boundary/Outlet/Outlet.cl",
                         "cfd/Boundary/Portal/Mirror.cl")
STEPS = [(INLET, "feed"), (INLET, "rates"), (OUTLET, "rates"), (OUTLET, "feed"), (PORTAL, "mirror"),
         (PORTAL, "unmirror"), (PORTAL, "teleport")]
# arrays each step writes (compared after it)
WRITES = {(INLET, "feed"): ("imove", "r", "u", "dudt", "rho", "drhodt", "m", "p"),
          (INLET, "rates"): ("u", "dudt", "drhodt"),
          (OUTLET, "rates"): ("u", "rho", "p", "dudt", "dudt_in", "drhodt", "drhodt_in"),
          (OUTLET, "feed"): ("imove", "r_in"),
          (PORTAL, "mirror"): ("r", "imirrored", "icell"),
          (PORTAL, "unmirror"): ("r",),
          (PORTAL, "teleport"): ("r",)}


def state(dims, seed=33):
    """A jittered dam break with nbuffer free rows at the end (imove = -255), planes cutting the fluid
    obliquely, every array a kernel reads filled with noise so that a skipped write shows."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    n, V = case["N"], (4 if dims == 3 else 2)
    nbuf = 97
    N = n + nbuf
    rng = np.random.default_rng(seed)

    def grow(a, fill):
        out = np.empty((N,) + a.shape[1:], a.dtype)
        out[:n] = a
        out[n:] = fill
        return out

    def vec(*xyz):
        a = np.zeros(V, np.float32)
        a[:len(xyz[:dims])] = xyz[:dims]
        return a

    def noise():
        a = rng.normal(size=(N, V)).astype(np.float32)
        if dims == 3:
            a[:, 3] = 0
        return a
    dmax = np.asarray(case["domain_max"], np.float32).ravel()[:V].copy()
    dmin = np.asarray(case["domain_min"], np.float32).ravel()[:V].copy()
    v = {"imove": grow(np.ascontiguousarray(case["imove"]), -255),
         "iset": grow(np.ascontiguousarray(case["iset"]).astype(np.uint32), 0),
         "r": grow(np.ascontiguousarray(case["r"]), dmax), "m": grow(np.ascontiguousarray(case["m"]), 0),
         "rho": grow(np.ascontiguousarray(case["rho"]), 1000.0), "u": noise(), "dudt": noise(), "dudt_in": noise(),
         "drhodt": rng.normal(size=N).astype(np.float32), "drhodt_in": rng.normal(size=N).astype(np.float32),
         "p": rng.normal(size=N).astype(np.float32), "refd": np.ascontiguousarray(case["refd"]),
         "imirrored": np.full(N, 7, np.int32), "icell": np.full(N, 0xFFFFFFFF, np.uint32)}
    v["r_in"] = v["r"].copy()
    fl = v["imove"] == 1
    rf = v["r"][fl][:, :dims]
    lo, hi, mid = rf.min(0), rf.max(0), np.median(rf, 0)
    nrm = vec(0.8, 0.6, 0.0)          # normalised, not axis aligned
    s = dict(N=N, nbuffer=nbuf, dt=1e-4, cs=float(case["cs"]), p0=250.0, g=vec(0.0, -9.81, 0.0) if dims == 2
             else vec(0.0, 0.0, -9.81), dr=float(case["dr"]), domain_max=dmax,
             inlet_r=vec(*mid), inlet_ru=vec(-0.06, 0.08, 0.0), inlet_rv=vec(0.0, 0.0, 0.11),
             inlet_N=np.array([7, 5] if dims == 3 else [11, 1], np.uint32), inlet_n=nrm, inlet_U=1.25,
             inlet_rFS=vec(*(hi + 0.01)), inlet_R=0.137, inlet_starving=1,
             outlet_r=vec(*(mid + 0.25 * (hi - mid))), outlet_n=nrm, outlet_U=0.75, outlet_rFS=vec(*(hi + 0.02)),
             portal_in_r=vec(*(lo + 0.2 * (hi - lo))), portal_out_r=vec(*(lo + 0.7 * (hi - lo))), portal_n=nrm,
             r_min=dmin - (dmax - dmin))     # (every mirrored position stays above it)
    # the link-list grid the portal re-hashes into (LinkList.cpp:423-447: cells of SUPPORT * h, 6 spare)
    L = 2.0 * np.float32(case["h"])
    nc = ((dmax[:dims] - s["r_min"][:dims]) / L).astype(np.uint32) + 6
    n_cells = np.ones(4, np.uint32)
    n_cells[:dims] = nc
    n_cells[3] = int(np.prod(n_cells[:3]))
    s["n_cells"] = n_cells
    v.update(s)
    return case, v


def args_of(v):
    """A private copy of the state (arrays copied, scalars shared)."""
    return {k: (x.copy() if isinstance(x, np.ndarray) and x.ndim and k not in SCALARS else x) for k, x in v.items()}


SCALARS = ("g", "domain_max", "inlet_r", "inlet_ru", "inlet_rv", "inlet_N", "inlet_n", "inlet_rFS", "outlet_r",
           "outlet_n", "outlet_rFS", "portal_in_r", "portal_out_r", "portal_n", "r_min", "n_cells", "refd")


def oracle_step(oracle, D, dims, key, b):
    """One step of STEPS on the oracle's restatement (oracle/aqo_kernels.c)."""
    c, N = oracle.call, b["N"]
    if key == (INLET, "feed"):
        c("inlet_feed", D, b["imove"], b["iset"], b["r"], b["u"], b["dudt"], b["rho"], b["drhodt"], b["m"], b["p"],
          b["refd"], N, b["nbuffer"], float(b["cs"]), float(b["p0"]), b["g"], float(b["dr"]), b["inlet_r"],
          b["inlet_ru"], b["inlet_rv"], b["inlet_N"], b["inlet_n"], float(b["inlet_U"]), b["inlet_rFS"],
          float(b["inlet_R"]), int(b["inlet_starving"]))
    elif key == (INLET, "rates"):
        c("inlet_rates", b["imove"], b["r"], b["u"], b["dudt"], b["drhodt"], N, b["inlet_r"], float(b["inlet_U"]),
          b["inlet_n"], dims)
    elif key == (OUTLET, "rates"):
        c("outlet_rates", b["imove"], b["iset"], b["r"], b["u"], b["rho"], b["p"], b["dudt"], b["dudt_in"],
          b["drhodt"], b["drhodt_in"], b["refd"], N, float(b["cs"]), float(b["p0"]), b["g"], b["outlet_r"],
          b["outlet_n"], float(b["outlet_U"]), b["outlet_rFS"], dims)
    elif key == (OUTLET, "feed"):
        c("outlet_feed", D, b["imove"], b["r_in"], N, b["domain_max"], b["outlet_r"], b["outlet_n"])
    elif key == (PORTAL, "mirror"):
        c("portal_mirror", D, b["r"], b["imirrored"], b["icell"], N, b["portal_in_r"], b["portal_out_r"],
          b["portal_n"], b["r_min"], b["n_cells"])
    elif key == (PORTAL, "unmirror"):
        c("portal_unmirror", b["r"], b["imirrored"], N, b["portal_in_r"], b["portal_out_r"], dims)
    elif key == (PORTAL, "teleport"):
        c("portal_teleport", b["r"], N, b["portal_in_r"], b["portal_out_r"], b["portal_n"], dims)
    else:
        raise KeyError(key)


def checks(v0, b, dims):
    """What the sequence must have done to the state (so that a kernel that does nothing fails)."""
    N, nbuf = b["N"], b["nbuffer"]
    nin = int(min(nbuf, int(b["inlet_N"][0]) * int(b["inlet_N"][1])))
    i0 = N - nbuf
    assert (b["m"][i0:i0 + nin] > 0).all() and (b["m"][i0 + nin:] == 0).all() and (b["imove"][i0 + nin:] == -255).all()
    gone = (b["imove"] == -256) & (v0["imove"] == 1)
    assert 0 < gone.sum() < (v0["imove"] == 1).sum()
    assert 0 < (b["imirrored"] == 1).sum() < N and set(np.unique(b["imirrored"])) == {0, 1}
    assert (b["icell"][b["imirrored"] == 1] < b["n_cells"][3]).all()
    assert (b["r"] != v0["r"]).any(1).sum() > nin      # teleported rows


# ---- the two pair sweeps of the portal (cfd/Boundary/Portal/Shepard.cl, Interactions.cl) --------------------
def portal_sweep_state(oracle, dims, n, hfac, seed=12):
    """A sorted dam break whose fluid is cut by an out portal (normal +x) three quarters along and an in portal
    one quarter along: Portal/Mirror.cl::mirror (the oracle's, bit-identical to the script) has moved the
    particles within a kernel support of the out plane to the in plane and re-hashed their cells -- the state
    the preset runs the portal sweeps on (cfd/portal.xml).  Outputs pre-filled with noise: the sweeps add."""
    import pipeline
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    N, V = s["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(seed)
    fl = s["imove"] == 1
    x = s["r"][fl][:, 0]
    lo, hi = float(x.min()), float(x.max())
    ctr = np.median(s["r"][fl], 0).astype(np.float32)
    pin, pout, pn = ctr.copy(), ctr.copy(), np.zeros(V, np.float32)
    pin[0], pout[0], pn[0] = lo + 0.25 * (hi - lo), lo + 0.75 * (hi - lo), 1.0
    # The planes lie more than four kernel supports apart, so that no cell a mirrored particle looks into holds a
    # mirrored particle: the reference's walk of a cell ends at the first row whose icell differs (BEGIN_NEIGHS),
    # i.e. AT a mirrored row that still sits in the run of its old cell, where the engines here go on to the
    # end of the 32-row tile -- a difference only a portal pair closer than that could show
    assert pout[0] - pin[0] > 8.5 * float(s["h"]), (pout[0] - pin[0]) / float(s["h"])
    D = oracle.make_defs(dims, s["h"])
    r = np.ascontiguousarray(s["r"]).copy()
    icell = np.ascontiguousarray(s["icell"]).copy()
    imirrored = np.full(N, 7, np.int32)
    rmin = np.zeros(V, np.float32)
    rmin[:len(s["r_min"][:V])] = np.asarray(s["r_min"], np.float32)[:V]
    oracle.call("portal_mirror", D, r, imirrored, icell, N, pin, pout, pn, rmin,
                np.ascontiguousarray(s["n_cells"], np.uint32))
    assert 0 < (imirrored[fl] == 1).sum() < fl.sum() and (icell < s["n_cells"][3]).all()
    u = np.zeros((N, V), np.float32)
    u[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)

    def noise_vec():
        a = np.zeros((N, V), np.float32)
        a[:, :dims] = rng.normal(size=(N, dims)).astype(np.float32)
        if dims == 3:
            a[:, 3] = 7.0
        return a
    x = dict(r=r, icell=icell, imirrored=imirrored, u=u, p=rng.uniform(-50.0, 300.0, N).astype(np.float32),
             grad_p=noise_vec(), lap_u=noise_vec(), div_u=rng.normal(size=N).astype(np.float32),
             shepard=rng.uniform(0.2, 0.9, N).astype(np.float32))
    return case, s, D, x


def portal_sweeps_oracle(oracle, s, D, x, morris):
    """The oracle's portal sweeps on a copy of the outputs; returns them."""
    import oracle.oracle as O
    L = O.make_ll(x["icell"], np.ascontiguousarray(s["ihoc"]), s["n_cells"], s["N"])
    w = {k: x[k].copy() for k in ("shepard", "grad_p", "lap_u", "div_u")}
    imove = np.ascontiguousarray(s["imove"])
    rho, m = np.ascontiguousarray(s["rho"]), np.ascontiguousarray(s["m"])
    oracle.call("portal_shepard", D, L, imove, x["imirrored"], x["r"], rho, m, w["shepard"])
    oracle.call("portal_interactions", D, L, imove, x["imirrored"], x["r"], x["u"], rho, m, x["p"], w["grad_p"],
                w["lap_u"], w["div_u"], int(morris))
    return w


def portal_rows(s, x, k):
    """The rows a portal sweep may change: mirrored fluid particles (Interactions.cl:65), mirrored particles
    of the classes -3 .. 1 for the Shepard factor (Shepard.cl:57)."""
    mv = np.asarray(s["imove"])
    if k == "shepard":
        return (x["imirrored"] == 1) & (mv >= -3) & (mv <= 1)
    return (x["imirrored"] == 1) & (mv == 1)
