"""Diagnostic: 2-GPU slabs vs 1 GPU, error of u / dudt against distance to the cut."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_mpi as T

if __name__ == "__main__":
    for maxiter in (1, 2):
        def patched(rank, size, port, q, n_total, steps, _m=maxiter):
            pass
        import types
        src = T._slab_rank
        n_total, steps = 40000, 2
        os.environ["AQ_DIAG_MAXITER"] = str(maxiter)
        one = T._run_slabs(1, n_total, steps)[0]
        two = T._run_slabs(2, n_total, steps)
        pos1 = {int(g): k for k, g in enumerate(one["fluid_index"])}
        for r in range(2):
            rows = np.array([pos1[int(g)] for g in two[r]["fluid_index"]])
            y = one["r"][rows][:, 1]
            cut = two[0]["slab"][1]
            d = np.abs(y - cut) / two[r]["h"]
            for k in ("u", "dudt", "rho"):
                a = one[k][rows].astype(np.float64); b = two[r][k].astype(np.float64)
                e = np.abs(a - b)
                e = e.max(1) if e.ndim > 1 else e
                sc = np.abs(a).max()
                line = " ".join("%.1e" % (e[(d >= lo) & (d < hi)].max() / sc if ((d >= lo) & (d < hi)).any() else 0)
                                for lo, hi in ((0, 2), (2, 4), (4, 8), (8, 1e9)))
                print("maxiter", maxiter, "rank", r, k, "rel err by distance/h [0-2,2-4,4-8,8+]:", line, flush=True)
