"""Regenerates tests/golden/time_scheme_outputs.npz: what the REFERENCE's own element-wise scripts of the hot path
(basic/time_scheme/{midpoint,euler,improved_euler}.cl and basic/Domain.cl, compiled behind oracle/ref_shim -- build
container only) make of seeded states; tests/test_oracle_golden.py holds the C restatements to these bits without
the reference tree.       python tests/golden/make_golden_elementwise.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402

DT, RELAX = 1e-3, 0.25
# (script, entry, outputs, restatement name, its arguments by variable name; "N" / "dt" / "relax" / "dims" literal)
SEQUENCE = (
    ("basic/time_scheme/midpoint.cl", "predictor", ("r_in", "u_in", "dudt_in", "rho_in", "drhodt_in"), "mp_predictor",
     ("r", "u", "dudt", "rho", "drhodt", "r_in", "u_in", "dudt_in", "rho_in", "drhodt_in", "N", "dims")),
    ("basic/time_scheme/midpoint.cl", "midpoint", ("u", "rho"), "mp_midpoint",
     ("imove", "u_in", "u", "dudt", "rho_in", "rho", "drhodt", "N", "dt", "dims")),
    ("basic/time_scheme/midpoint.cl", "relax", ("dudt", "drhodt"), "mp_relax",
     ("imove", "dudt_in", "dudt", "drhodt_in", "drhodt", "N", "relax", "dims")),
    ("basic/time_scheme/midpoint.cl", "corrector", ("r", "u", "rho"), "mp_corrector",
     ("imove", "r_in", "r", "u_in", "u", "dudt", "rho_in", "rho", "drhodt", "N", "dt", "dims")),
    ("basic/time_scheme/euler.cl", "corrector", ("r", "u", "rho"), "euler_corrector",
     ("imove", "r", "u", "dudt", "rho", "drhodt", "N", "dt", "dims")),
    ("basic/time_scheme/improved_euler.cl", "predictor", ("r_in", "u_in", "dudt_in", "rho_in", "drhodt_in"), "ie_predictor",
     ("imove", "r", "u", "dudt", "rho", "drhodt", "r_in", "u_in", "dudt_in", "rho_in", "drhodt_in", "N", "dt", "dims")),
    ("basic/time_scheme/improved_euler.cl", "corrector", ("r", "u", "rho"), "ie_corrector",
     ("imove", "r", "u", "dudt", "rho", "drhodt", "dudt_in", "drhodt_in", "N", "dt", "dims")),
    (None, "spoil", (), None, ()),   # a seventh of the positions far outside the box and one NaN, right in front of:
    ("basic/Domain.cl", "entry", ("imove", "r_in", "u_in", "dudt_in", "m"), "domain",
     ("imove", "r_in", "u_in", "dudt_in", "m", "N", "domain_min", "domain_max", "dims")),
)


def state(dims, seed=5):
    """A dam break's particle classes with random fields (spoil() drives the removal branch of
    basic/Domain.cl:48-90)."""
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(seed)
    v = {k: np.ascontiguousarray(case[k]).copy() for k in ("imove", "iset", "id", "r", "u", "dudt", "rho", "drhodt", "m",
                                                                 "normal", "tangent")}
    for k in ("r", "u", "dudt"):
        v[k + "_in"] = rng.normal(size=(N, V)).astype(np.float32)
    for k in ("rho", "drhodt"):
        v[k + "_in"] = rng.normal(size=N).astype(np.float32)
    v.update(N=N, dt=DT, relax_midpoint=RELAX, domain_min=case["domain_min"], domain_max=case["domain_max"])
    return v, N


def spoil(v):
    v["r_in"][::7] *= 100.0
    v["r_in"][3, 0] = np.nan


def main():
    from oracle import ref
    fx = {}
    for dims in (2, 3):
        v, N = state(dims)
        R = ref.Ref(dims, 0.1)
        for script, entry, outs, _, _ in SEQUENCE:
            if script is None:
                spoil(v)
                continue
            R.run(script, entry, N, v)
            for k in outs:
                fx["%dD_%s_%s_%s" % (dims, os.path.basename(script)[:-3], entry, k)] = v[k].copy()
    np.savez_compressed(os.path.join(HERE, "time_scheme_outputs.npz"), **fx)
    print(len(fx), "arrays,", os.path.getsize(os.path.join(HERE, "time_scheme_outputs.npz")), "bytes on disk")


if __name__ == "__main__":
    main()
