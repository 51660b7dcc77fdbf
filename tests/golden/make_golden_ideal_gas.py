"""Regenerates tests/golden/ideal_gas_outputs.npz: what the REFERENCE's own scripts of resources/Scripts/cfd/ideal_gas
(compiled behind oracle/ref_shim: build container only, needs /root/reference) make of the seeded inputs of
tests/test_oracle_vs_reference.py::_ideal_gas_state / riemann_inputs.  tests/test_oracle_golden.py compares the C
restatement with these values where neither the reference tree nor oracle/_ref exists.

    python tests/golden/make_golden_ideal_gas.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import pipeline  # noqa: E402
from oracle import ref  # noqa: E402
from test_oracle_vs_reference import _ideal_gas_state, riemann_inputs  # noqa: E402

SEQUENCE = (("EOS.cl", "entry"), ("Rates.cl", "entry"), ("TimeStep.cl", "entry"),
            ("time_scheme/midpoint.cl", "predictor"), ("riemann/Rates.cl", "entry"),
            ("time_scheme/midpoint.cl", "midpoint"), ("time_scheme/midpoint.cl", "relax"),
            ("time_scheme/midpoint.cl", "corrector"))
OUTPUTS = ("p", "deintdt", "dt_var", "eint", "eint_in", "deintdt_in")


def main():
    fx = {}
    for dims in (2, 3):
        a, _, N = _ideal_gas_state(dims, 13)
        R = ref.Ref(dims, a["h"])
        for script, entry in SEQUENCE:
            R.run("cfd/ideal_gas/" + script, entry, N, a)
        a["eint_in"][...] = a["eint"]
        R.run("cfd/ideal_gas/Sort.cl", "entry", N, a)
        R.run("cfd/ideal_gas/time_scheme/euler.cl", "predictor", N, a)
        R.run("cfd/ideal_gas/time_scheme/euler.cl", "corrector", N, a)
        a["deintdt"][...] = a["work_density"]
        R.run("cfd/ideal_gas/time_scheme/improved_euler.cl", "corrector", N, a)
        R.run("cfd/ideal_gas/time_scheme/improved_euler.cl", "predictor", N, a)
        R.run("cfd/ideal_gas/symmetry/Mirror.cl", "set", N, a)
        for k in OUTPUTS:
            fx["elementwise_%dD_%s" % (dims, k)] = a[k]
    for dims, n, hfac in ((2, 40, 3.0), (3, 10, 2.0)):
        case = cases.dam_break(dims, n, hfac)
        s = pipeline.oracle_linklist_and_sort(case)
        x = riemann_inputs(s)
        c = pipeline.RefState(ref.Ref(dims, case["h"]), s)
        for k in ("u", "p", "iset", "grad_p", "div_u"):
            c.set(k, x[k])
        c.v["gamma"] = x["gamma"].copy()
        c.v["work_density"] = x["work_density"].copy()
        c.run("cfd/ideal_gas/riemann/Interactions.cl")
        for k in ("grad_p", "div_u", "work_density"):
            fx["riemann_%dD_%s" % (dims, k)] = c.get(k)
    np.savez_compressed(os.path.join(HERE, "ideal_gas_outputs.npz"), **fx)
    for k, v in fx.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
