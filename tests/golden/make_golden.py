"""Regenerates tests/golden/*.npz from the reference's own test inputs.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference ships no value-level golden outputs for the link-list / sort
(SURVEY.md section 8c): its tests assert PROPERTIES of the outputs on fixed
random inputs (tests/{2D,3D}/LinkList/cMake/check.py, RadixSort/cMake/check.py,
2D/MPI_plane/cMake/check.py).  So the fixtures hold the reference's INPUTS
(particles.dat, float64 text -> float32 exactly as FastASCII.cpp:46-120 parses
them with strtod + narrowing) and the tests re-assert the reference's own
property checks on our outputs.
"""
import os
import numpy as np

REF = "/root/reference/tests"
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            rows.append([float(t) for t in line.replace(",", " ").split()])
    return np.array(rows, dtype=np.float64)


def main():
    fx = {}
    for d in ("2D", "3D"):
        a = load(f"{REF}/{d}/LinkList/cMake/particles.dat")
        fx[f"linklist_{d}_r"] = a.astype(np.float32)
        b = load(f"{REF}/{d}/RadixSort/cMake/particles.dat")
        nv = 2 if d == "2D" else 4
        fx[f"radixsort_{d}_r"] = b[:, :nv].astype(np.float32)
        fx[f"radixsort_{d}_f"] = b[:, nv].astype(np.uint32)
    p = load(f"{REF}/2D/MPI_plane/cMake/particles.dat")
    # columns: r(2) u(2) dudt(2) rho drhodt m imove (main_serial.xml:26)
    fx["mpi_plane_2D"] = p.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "reference_inputs.npz"), **fx)
    for k, v in fx.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
