"""Regenerates tests/golden/hotpath_sweeps_outputs.npz: what the REFERENCE's own hot-path scripts (compiled behind
oracle/ref_shim: build container only, needs /root/reference) make of the cell-sorted dam-break states of
tests/test_oracle_vs_reference.py -- the sequence of tests/pipeline.py::named_sweeps (EOS, MLS, Shepard,
Interactions, sensors, delta-SPH full / lapp / full_mls / lapp_corr, BIe interactions / rates / p_boundary /
force_press / ElasticBounce / PST, Rates, residuals, TimeStep).  tests/test_oracle_golden.py holds the C restatement to these bits where neither the reference
tree nor oracle/_ref exists.

    python tests/golden/make_golden_hotpath.py
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

SWEEP_CASES = ((3, 10, 3.0), (2, 40, 4.0))    # (dims, n, hfac): two of the states tests/test_gpu_kernels.py runs on the GPU


def main():
    fx = {}
    for dims, n, hfac in SWEEP_CASES:
        case = cases.dam_break(dims, n, hfac)
        s = pipeline.oracle_linklist_and_sort(case)
        out = pipeline.ref_sweeps(ref.Ref(dims, case["h"]), s)
        for k, v in out.items():
            fx["sweeps_%dD_%s" % (dims, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "hotpath_sweeps_outputs.npz"), **fx)
    print(len(fx), "arrays,", sum(v.nbytes for v in fx.values()), "bytes raw,",
          os.path.getsize(os.path.join(HERE, "hotpath_sweeps_outputs.npz")), "bytes on disk")


if __name__ == "__main__":
    main()
