"""Outputs of the reference's OWN open-boundary scripts (cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl,
Portal/Mirror.cl, compiled behind oracle/ref_shim where /root/reference is) on the state and in the order
of tests/open_boundary_common.py: every array a kernel writes, after that kernel.

    python tests/golden/make_golden_open_boundary.py      ->  tests/golden/open_boundary_outputs.npz

tests/test_oracle_golden.py holds the oracle's restatement to these bits and tests/test_zz_gpu_open_boundary.py
the CUDA kernels, without the reference tree."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import open_boundary_common as ob  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    assert ref.build(), "needs /root/reference"
    out = {}
    for dims in (2, 3):
        case, v = ob.state(dims)
        R = ref.Ref(dims, case["h"])
        a = ob.args_of(v)
        for step, key in enumerate(ob.STEPS):
            R.run(key[0], key[1], v["N"], a)
            for k in ob.WRITES[key]:
                out["%dD_step%d_%s" % (dims, step, k)] = a[k].copy()
    np.savez_compressed(os.path.join(HERE, "open_boundary_outputs.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
