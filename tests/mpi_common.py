"""Helpers of the multi-device tests: the reference's tests/2D/MPI_plane case run
serially and on two ranks by the CPU oracle interpreter."""
import threading

import numpy as np

from aquagpusph_b200 import cases, casegen
from oracle import interp

FIELDS = cases.MPI_PLANE_FIELDS


def oracle_serial(table, steps=1, overrides=None):
    arrays, _ = cases.mpi_plane(table)
    I = interp.Interpreter(casegen.instantiate_plain("mpi_plane_2d_serial", overrides), 2)
    for k, a in arrays.items():
        I.V[k][...] = a
    for _ in range(steps):
        I.step()
    return {k: I.unsorted(k) for k in FIELDS}


def mpi_xml(overrides=None, fixed_mask=False):
    txt = casegen.instantiate_plain("mpi_plane_2d_mpi", overrides)
    return casegen.halo_mask_after_sort(txt) if fixed_mask else txt


def oracle_rank(table, rank, transport, steps=1, overrides=None, out=None, fixed_mask=False):
    arrays, own = cases.mpi_plane(table, rank)
    I = interp.Interpreter(mpi_xml(overrides, fixed_mask), 2, rank=rank, size=2, transport=transport)
    for k, a in arrays.items():
        I.V[k][...] = a
    for _ in range(steps):
        I.step()
    res = {k: I.unsorted(k) for k in FIELDS}
    res["own"] = own
    if out is not None:
        out[rank] = res
    return res


def oracle_two_ranks_threads(table, steps=1, overrides=None, fixed_mask=False):
    tr = interp.LocalTransport(2)
    out, errs = {}, []

    def work(rank):
        try:
            oracle_rank(table, rank, tr, steps, overrides, out, fixed_mask)
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            tr._barrier.abort()
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return out


def check_against_serial(serial, ranks, tol=1e-6, relative=False):
    """tests/2D/MPI_plane/cMake/check.py: every particle of the serial run must be
    found, in order, in the output of the process that owns it, every field within
    `tol` (absolute, or relative to the field's maximum); buffer rows are ignored."""
    worst = 0.0
    for rank, res in ranks.items():
        own = res["own"]
        n = len(own)
        for k in FIELDS:
            a = np.asarray(serial[k])[own].astype(np.float64)
            b = np.asarray(res[k])[:n].astype(np.float64)
            err = np.abs(a - b).max() if a.size else 0.0
            if relative:
                err /= max(np.abs(a).max(), 1e-30)
            worst = max(worst, err)
            assert err <= tol, "rank %d field %s: max err %.3e" % (rank, k, err)
    return worst
