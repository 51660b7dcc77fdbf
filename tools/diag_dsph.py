"""Diagnostic: the delta-SPH slab pipeline on N GPUs against the 116-tool pipeline on one GPU, step by
step: which live fluid particle has no twin (position off), on which rank, after which step.
    python tools/diag_dsph.py [size] [n_total] [max_steps]      (env switches are inherited by the ranks)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_mpi as T

if __name__ == "__main__":
    import torch.multiprocessing as mp
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    kw = dict(seed=5, jitter=0.45, uscale=2.0, iter_midpoint_max=3)
    only_ = os.environ.get("AQ_DIAG_ONLY")
    tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith(("AQC_", "AQUA_")))
    only = int(os.environ.pop("AQ_DIAG_ONLY", "0"))
    for steps in ([only] if only else range(1, max_steps + 1)):
        mpx = mp.get_context("spawn")
        q = mpx.Queue()
        p = mpx.Process(target=T._single_116_rank, args=(q, n_total, steps, kw))
        p.start()
        one = T._collect([p], q, 1, 600)[0]
        many = T._run_slabs(size, n_total, steps, whole=True, delta_sph=True, **kw)
        fl1, matches = T._match_rows(one, [many[r] for r in range(size)])
        allm = np.concatenate([m[1] for m in matches])
        uniq, cnt = np.unique(allm, return_counts=True)
        missing = np.setdiff1d(fl1, uniq)
        print("[%s] steps %d: live %s of %d, unique twins %d, twice-matched %s, unmatched 1-GPU rows %s"
              % (tag, steps, [len(m[0]) for m in matches], len(fl1), len(uniq), uniq[cnt > 1].tolist(),
                 missing.tolist()), flush=True)
        for r in range(size):
            rows, rows1, d = matches[r]
            bad = np.flatnonzero(d > 1e-4 * one["h"])
            print("  rank %d: max twin distance %.3e (h %.4f), %d rows beyond 1e-4 h, slab %s"
                  % (r, d.max(), one["h"], len(bad), many[r]["slab"]), flush=True)
            for b in bad[:6]:
                print("    row %d r=%s u=%s rho=%.5f | nearest 1-GPU row %d r=%s u=%s (d %.3e)"
                      % (rows[b], many[r]["r"][rows[b]][:3], many[r]["u"][rows[b]][:3], many[r]["rho"][rows[b]],
                         rows1[b], one["r"][rows1[b]][:3], one["u"][rows1[b]][:3], d[b]), flush=True)
            for k in ("r", "u", "rho", "dudt"):
                a = one[k][rows1].astype(np.float64)
                b_ = many[r][k][rows].astype(np.float64)
                e = np.abs(a - b_)
                e = e.max(1) if e.ndim > 1 else e
                good = d <= 1e-4 * one["h"]
                print("    %-5s rel err (twins found) %.3e" % (k, e[good].max() / max(np.abs(one[k][fl1]).max(), 1e-30)),
                      flush=True)
        for m_ in missing[:6]:
            print("  unmatched 1-GPU row %d r=%s u=%s" % (m_, one["r"][m_][:3], one["u"][m_][:3]), flush=True)
