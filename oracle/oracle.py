"""ctypes front-end of the CPU oracle (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  It loads oracle/liboracle.so (the plain-C
restatement, built by `make -C oracle`) and, when present, the reference's own
OpenCL sources compiled behind the shim (oracle/_ref/libaquaref{2,3}d.so).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Defs(C.Structure):
    _fields_ = [("dims", C.c_int), ("H", C.c_float), ("CONW", C.c_float),
                ("CONF", C.c_float), ("SUPPORT", C.c_float)]


class LL(C.Structure):
    _fields_ = [("icell", C.c_void_p), ("ihoc", C.c_void_p),
                ("ncells", C.c_uint32 * 4), ("N", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.aqo_define_round6.restype = C.c_float
        _LIB.aqo_define_round6.argtypes = [C.c_float]
        for n in ("aqo_reduce_min", "aqo_reduce_max", "aqo_reduce_sum_tree"):
            getattr(_LIB, n).restype = C.c_float
        _LIB.aqo_reduce_max_u32.restype = C.c_uint32
    return _LIB


def _arg(a):
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, (float, np.floating)):
        return C.c_float(float(a))
    if isinstance(a, (bool, int, np.integer)):
        return C.c_int64(int(a)) if int(a) > 0x7FFFFFFF else C.c_int32(int(a) if int(a) < 2**31 else int(a) - 2**32)
    if isinstance(a, C.Structure):
        return C.byref(a)
    if a is None:
        return C.c_void_p(0)
    return a


def ptrs(arrays):
    """An array of pointers (float* const*) to the given contiguous numpy arrays."""
    arrays = list(arrays)
    assert all(isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] for a in arrays)
    p = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    p._keep = arrays
    return p


def call(name, *args):
    """Call aqo_<name> converting numpy arrays / python scalars."""
    return getattr(lib(), "aqo_" + name)(*[_arg(a) for a in args])


_POOL = None
THREADS = 1


def set_threads(n):
    """Spread the particle loops of `pcall` kernels over n host threads."""
    global _POOL, THREADS
    from concurrent.futures import ThreadPoolExecutor
    THREADS = max(1, int(n))
    _POOL = ThreadPoolExecutor(THREADS) if THREADS > 1 else None


def pcall(name, N, *args):
    """call() with the rows [0, N) split across the thread pool (ctypes drops the
    GIL; every oracle kernel only writes row i inside its i-loop)."""
    if _POOL is None or N < 4096:
        return call(name, *args)
    L = lib()
    fn = getattr(L, "aqo_" + name)
    cargs = [_arg(a) for a in args]
    # more chunks than threads: sorted particle sets are far from uniform in cost
    chunks = THREADS * 8
    step = (N + chunks - 1) // chunks

    def work(k):
        L.aqo_set_range(C.c_uint32(k * step), C.c_uint32(min(N, (k + 1) * step)))
        try:
            fn(*cargs)
        finally:
            L.aqo_set_range(C.c_uint32(0), C.c_uint32(0xFFFFFFFF))

    list(_POOL.map(work, range(chunks)))


def vs(dims):
    return 4 if dims == 3 else 2


def ms(dims):
    return 16 if dims == 3 else 4


def make_defs(dims, h):
    d = Defs()
    lib().aqo_make_defs(C.byref(d), C.c_int(dims), C.c_float(h))
    return d


def make_ll(icell, ihoc, ncells, N):
    ll = LL()
    ll.icell = icell.ctypes.data
    ll.ihoc = ihoc.ctypes.data
    for k in range(4):
        ll.ncells[k] = int(ncells[k])
    ll.N = int(N)
    ll._keep = (icell, ihoc)
    return ll


def linklist(r, dims, support, h, rmin=None, rmax=None, recompute=True):
    """LinkList::_execute (LinkList.cpp:326-494) on host arrays.

    Returns dict(rmin, rmax, ncells, icell (sorted), ihoc, perm=id_unsorted,
    inv_perm=id_sorted)."""
    r = np.ascontiguousarray(r, dtype=np.float32)
    N = r.shape[0]
    rmin = np.zeros(vs(dims), np.float32) if rmin is None else np.array(rmin, np.float32)
    rmax = np.zeros(vs(dims), np.float32) if rmax is None else np.array(rmax, np.float32)
    if recompute:
        call("minmax", r, N, dims, rmin, rmax)
    ncells = np.zeros(4, np.uint32)
    rc = call("ncells", rmin, rmax, dims, float(support), float(h), ncells)
    if rc:
        raise RuntimeError("Invalid number of cells")
    icell = np.zeros(N, np.uint32)
    ihoc = np.zeros(int(ncells[3]), np.uint32)
    perm = np.zeros(N, np.uint32)
    inv = np.zeros(N, np.uint32)
    rc = call("linklist", r, N, dims, float(support), float(h), 0, rmin, rmax,
              ncells, icell, ihoc, C.c_size_t(ihoc.size), perm, inv)
    assert rc == 0
    return dict(rmin=rmin, rmax=rmax, ncells=ncells, icell=icell, ihoc=ihoc,
                perm=perm, inv_perm=inv)


def scatter(src, idx):
    """out[idx[i]] = src[i] (basic/Sort.cl, UnSort.cl.in)."""
    src = np.ascontiguousarray(src)
    out = np.empty_like(src)
    eb = src.dtype.itemsize * (int(np.prod(src.shape[1:])) if src.ndim > 1 else 1)
    call("scatter", out, src, np.ascontiguousarray(idx, np.uint32),
         src.shape[0], C.c_size_t(eb))
    return out
