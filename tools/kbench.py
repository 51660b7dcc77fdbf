"""Per-kernel timing of the hot-path kernels on a sorted dam-break state.

    python tools/kbench.py [--n 1000000] [--reps 5] [--only NAME] [--hfac 3]

Prints one JSON line per kernel (CUDA-event time, algorithmic bytes, pair counts).
Used under gpurun / ncu; not part of the product."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import _lib, cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--only", default="")
    ap.add_argument("--hfac", type=float, default=3.0)
    ap.add_argument("--case", default="spheric2")
    ap.add_argument("--cache", type=int, default=0, help="1: pair-mask cache on (sweeps read the masks)")
    ap.add_argument("--suborder", type=int, default=0,
                    help="input pre-ordered by S x S x S sub-cells inside every link-list cell (-1: random order)")
    a = ap.parse_args()
    t0 = time.time()
    if a.case == "spheric2":
        c = cases.spheric2_dam_break(a.n, a.hfac, seed=1)
    else:
        c = cases.lattice(int(round(a.n ** (1 / 3))), a.hfac)
    N, dims = c["N"], c["dims"]
    if a.suborder:
        # the link-list sort is stable: the order inside a cell is the input order
        r = c["r"][:, :dims].astype(np.float64)
        t = (r - r.min(0)) / (2.0 * c["h"])
        if a.suborder > 0:
            sub = np.minimum((np.modf(t)[0] * a.suborder).astype(np.int64), a.suborder - 1)
            key = sub[:, 0] + a.suborder * sub[:, 1] + (a.suborder ** 2 * sub[:, 2] if dims == 3 else 0)
            order = np.argsort(key, kind="stable")
        else:
            order = np.random.default_rng(7).permutation(N)
        for k, x in list(c.items()):
            if isinstance(x, np.ndarray) and x.ndim >= 1 and x.shape[0] == N and k not in ("refd", "visc_dyn", "delta"):
                c[k] = np.ascontiguousarray(x[order])
    ctx = _lib.Context(0, dims=dims, h=c["h"])
    if a.cache:
        ctx.pairs_cache(True)
    V, M = (4 if dims == 3 else 2), (16 if dims == 3 else 4)
    v = {k: ctx.array(c[k]) for k in ("id", "iset", "imove", "r", "normal", "tangent", "rho", "m",
                                      "u", "dudt", "drhodt", "refd", "visc_dyn", "delta")}
    for k in ("id", "iset", "imove", "r", "normal", "tangent", "rho", "m", "u", "dudt", "drhodt"):
        v[k + "_in"] = ctx.empty(c[k].shape, c[k].dtype)
    for k in ("binormal", "grad_p", "lap_u", "lap_p_corr", "grad_w_bi", "r_bak", "r_in2"):
        v[k] = ctx.zeros((N, V), np.float32)
    for k in ("p", "div_u", "shepard", "lap_p", "div_u_bi", "dt_var", "residual_midpoint"):
        v[k] = ctx.zeros(N, np.float32)
    v["mls"] = ctx.zeros((N, M), np.float32)
    v["n_neighs"] = ctx.zeros(N, np.uint32)
    for k in ("icell", "id_sorted", "id_unsorted"):
        v[k] = ctx.empty(N, np.uint32)
    for k in ("N", "cs", "p0", "g", "courant", "dt_Ma", "dt_min", "h", "domain_min", "domain_max"):
        v[k] = c[k]
    v.update(dt=1e-4, neighs_limit=100000, mls_imove=1, relax_midpoint=0.1)
    ihoc = None

    def linklist():
        nonlocal ihoc
        _, _, nc, ihoc = ctx.linklist(v["r_in"], 2.0, c["h"], v["icell"], ihoc, v["id_unsorted"],
                                      v["id_sorted"])
        v["ihoc"], v["n_cells"] = ihoc, nc

    def backup():
        for k in ("id", "iset", "imove", "normal", "tangent", "m"):
            ctx.copy(v[k + "_in"], v[k])

    K = lambda s, e="entry": (lambda: ctx.launch(s, e, v))  # noqa: E731
    # prepare: predictor, link-list, sort, EOS
    K("basic/time_scheme/midpoint.cl", "predictor")()
    linklist(); backup(); K("basic/Sort.cl", "stage1")(); K("basic/Sort.cl", "stage2")()
    ctx.copy(v["dudt"], v["dudt_in"]); ctx.copy(v["drhodt"], v["drhodt_in"])
    K("basic/EOS.cl")()
    ctx.sync()
    nc = v["n_cells"]
    imv = v["imove"].get()
    nfl = int((imv == 1).sum())
    print(json.dumps(dict(case=a.case, N=N, n_fluid=nfl, n_cells=[int(x) for x in nc], h=c["h"],
                          setup_s=round(time.time() - t0, 2), sm=ctx.sm_count())), flush=True)

    def relink():
        # state is already sorted: re-running measures the steady-state (nearly sorted) case
        K("basic/time_scheme/midpoint.cl", "predictor")()
        linklist()

    vb = 16 if dims == 3 else 8
    tests = [
        ("linklist", relink, (vb + 20 + 4 + 16 * 2 + 8 + 4) * N),
        # the link-list build alone on the cell-ordered state: SURVEY 8(d)'s 84 B + 4 n_cells.w / N per particle
        ("linklist_only", linklist, 84 * N + 4 * int(nc[3])),
        ("sort_stage1+2", lambda: (backup(), K("basic/Sort.cl", "stage1")(), K("basic/Sort.cl", "stage2")()), 212 * N),
        ("predictor", K("basic/time_scheme/midpoint.cl", "predictor"), 112 * N),
        ("eos", K("basic/EOS.cl"), 16 * N),
        ("neighs", K("basic/neighs.cl"), 8 * N),
        ("interactions", K("cfd/Interactions.cl"), 88 * N),
        ("shepard", K("cfd/Shepard.cl"), 36 * N),
        ("lapp", K("cfd/deltaSPH.cl", "lapp"), 40 * N),
        ("full", K("cfd/deltaSPH.cl", "full"), 52 * N),
        ("lapp_corr", K("cfd/deltaSPH.cl", "lapp_corr"), 60 * N),
        ("fused_fluid", lambda: ctx.launch_fused([("cfd/Shepard.cl", "entry"), ("cfd/Interactions.cl", "entry"),
                                                  ("cfd/deltaSPH.cl", "full"), ("cfd/deltaSPH.cl", "lapp")], v), 112 * N),
        ("mls", K("basic/MLS.cl"), 92 * N),
        # cache only: the masks are dropped before every launch (time = builder + reading sweep)
        # (enable() also clears the pay-off history, which would otherwise suspend these builds)
        ("build+shepard", lambda: (ctx.pairs_cache(True) if a.cache else None, K("cfd/Shepard.cl")()), 36 * N),
        ("bie_interactions", K("cfd/Boundary/BIe/Interactions.cl"), 76 * N),
        ("bie_p_boundary", K("cfd/Boundary/BIe/Interactions.cl", "p_boundary"), 36 * N),
        ("bie_elastic_bounce", K("cfd/Boundary/BIe/ElasticBounce.cl"), 68 * N),
        ("bie_pst", K("cfd/Boundary/BIe/PST.cl"), 60 * N),
        ("rates", K("cfd/Rates.cl"), 64 * N),
        ("corrector", K("basic/time_scheme/midpoint.cl", "corrector"), 96 * N),
        ("timestep", K("cfd/TimeStep.cl"), 24 * N),
        ("reduce_min", lambda: ctx.reduce(_lib.OP_MIN, v["dt_var"], host=False), 4 * N),
    ]
    for name, fn, nbytes in tests:
        if a.only and name not in a.only.split(","):
            continue
        if name == "corrector":
            ctx.copy(v["r_bak"], v["r"])
        for _ in range(a.warm):
            fn()
        prof = getattr(_lib.lib(), "aqc_debug_s3prof", None) if hasattr(_lib.lib(), "aqc_debug_s3prof") else None
        if prof is not None:
            prof(None, 1)
        e0, e1 = ctx.event(), ctx.event()
        l0 = ctx.launch_count()
        ctx.record(e0)
        for _ in range(a.reps):
            fn()
        ctx.record(e1)
        ms = ctx.elapsed_ms(e0, e1) / a.reps
        if name == "corrector":
            ctx.copy(v["r"], v["r_bak"])
        if prof is not None:
            import ctypes
            buf = (ctypes.c_ulonglong * 32)()
            prof(buf, 1)
            if buf[31]:
                names = ["wait_full", "filter", "balanced", "release", "tail", "loop", "endsync", "-"]
                tot = float(sum(buf[0:7])) or 1.0
                toti = float(sum(buf[8:15])) or 1.0
                print(json.dumps(dict(s3prof=name,
                    working={n: round(buf[k] / tot, 3) for k, n in enumerate(names[:7])},
                    working_Mcycles=round(tot / 1e6 / a.reps, 1),
                    idle_warps={n: round(buf[8 + k] / toti, 3) for k, n in enumerate(names[:7])},
                    idle_Mcycles=round(toti / 1e6 / a.reps, 1),
                    producer=dict(wait_empty=buf[16], stage=buf[17], loop=buf[20], tail=buf[28], endsync=buf[30 - 0] * 0 + buf[30 - 0] * 0),
                    producer_Mcycles=round((buf[16] + buf[17] + buf[20]) / 1e6 / a.reps, 1),
                    rounds_per_pass=round(buf[29] / buf[31], 1), passes=buf[31] // a.reps)), flush=True)
        print(json.dumps(dict(kernel=name, ms=round(ms, 4), launches=(ctx.launch_count() - l0) // a.reps,
                              alg_GBs=round(nbytes / ms / 1e6, 1),
                              Mparticles_s=round(N / ms / 1e3, 1))), flush=True)
    if a.cache:
        print(json.dumps(dict(pairs_cache=ctx.pairs_cache_stats())), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
