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
