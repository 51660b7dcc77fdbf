"""Regenerates tests/golden/linklist_tool_outputs.npz: the cell index of every particle and the head-of-cell table
as the REFERENCE's own tool-layer kernels compute them (aquagpusph/CalcServer/LinkList.cl.in: iCell :54-85, iHoc
:32-42, linkList :92-113, compiled behind oracle/ref_shim -- build container only), on the reference's LinkList test
particles (tests/{2D,3D}/LinkList/cMake/particles.dat, in tests/golden/reference_inputs.npz), on a dam break and on
seeded random positions.  iCell runs on the UNSORTED positions; the sorted order it is fed back in for linkList is
numpy's stable argsort of those cells (what RadixSort's property tests demand of any stable sort).

    python tests/golden/make_golden_linklist.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle, ref  # noqa: E402


def inputs(dims):
    """[(name, r (N, V) float32, h)]: shared with tests/test_oracle_golden.py"""
    V = 4 if dims == 3 else 2
    g = np.load(os.path.join(HERE, "reference_inputs.npz"))
    rng = np.random.default_rng(17)
    case = cases.dam_break(dims, 12 if dims == 3 else 50, 2.0)
    rnd = np.zeros((5000, V), np.float32)
    rnd[:, :dims] = rng.normal(size=(5000, dims)).astype(np.float32) * 3.0
    out = []
    for name, r, h in (("reference", g["linklist_%dD_r" % dims], 0.1), ("dambreak", case["r"], case["h"]),
                       ("random", rnd, 0.37)):
        r = np.ascontiguousarray(r, np.float32)
        if r.shape[1] != V:
            rr = np.zeros((r.shape[0], V), np.float32)
            rr[:, :min(V, r.shape[1])] = r[:, :min(V, r.shape[1])]
            r = rr
        out.append((name, r, float(h)))
    return out


def main():
    fx = {}
    for dims in (2, 3):
        R = ref.Ref(dims, 0.1)
        for name, r, h in inputs(dims):
            N = r.shape[0]
            ll = oracle.linklist(r, dims, 2.0, h)       # (r_min / n_cells: host arithmetic, LinkList.cpp:185-232)
            icell = np.zeros(N, np.uint32)
            R.run("LinkList.cl", "iCell", N, dict(icell=icell, r=r, N=N, r_min=ll["rmin"], support=2.0, h=h,
                                                   n_cells=ll["ncells"]))
            perm = np.argsort(icell, kind="stable").astype(np.uint32)
            sorted_cells = np.ascontiguousarray(icell[perm])
            ncw = int(ll["ncells"][3])
            ihoc = np.zeros(ncw, np.uint32)
            R.run("LinkList.cl", "iHoc", ncw, dict(ihoc=ihoc, N=N, n_cells=ll["ncells"]))
            R.run("LinkList.cl", "linkList", N, dict(icell=sorted_cells, ihoc=ihoc, N=N))
            key = "%dD_%s" % (dims, name)
            fx[key + "_icell_unsorted"] = icell
            fx[key + "_ihoc"] = ihoc
            fx[key + "_ncells"] = np.asarray(ll["ncells"], np.uint32)
    np.savez_compressed(os.path.join(HERE, "linklist_tool_outputs.npz"), **fx)
    print(len(fx), "arrays,", os.path.getsize(os.path.join(HERE, "linklist_tool_outputs.npz")), "bytes on disk")


if __name__ == "__main__":
    main()
