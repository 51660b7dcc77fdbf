"""GPU parity: LinkList / RadixSort / UnSort tools through the C-ABI vs the oracle.
Bit-exact (integer / index work)."""
import numpy as np
import pytest

import cases
from aquagpusph_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx3():
    c = _lib.Context(0, dims=3, h=0.1)
    yield c
    c.close()


def run_linklist(dims, r, support, h):
    ctx = _lib.Context(0, dims=dims, h=h)
    N = r.shape[0]
    d_r = ctx.array(r)
    icell, perm, inv = (ctx.empty(N, np.uint32) for _ in range(3))
    rmin, rmax, nc, ihoc = ctx.linklist(d_r, support, h, icell, None, perm, inv)
    out = dict(rmin=rmin, rmax=rmax, ncells=nc, icell=icell.get(), ihoc=ihoc.get()[:nc[3]],
               perm=perm.get(), inv_perm=inv.get())
    ctx.close()
    return out


def assert_ll_equal(a, b):
    for k in ("rmin", "rmax", "ncells", "icell", "ihoc", "perm", "inv_perm"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


@pytest.mark.parametrize("d", ["2D", "3D"])
def test_linklist_reference_inputs(golden, oracle, d):
    """The reference's tests/{2D,3D}/LinkList input (isolated, coincident and
    far-away particles), h = 0.1 as in its main.xml; two sets concatenated."""
    dims = int(d[0])
    r1 = golden[f"linklist_{d}_r"]
    r = np.concatenate([r1, r1])
    ref = oracle.linklist(r, dims, 2.0, 0.1)
    got = run_linklist(dims, r, 2.0, 0.1)
    assert_ll_equal(ref, got)
    # the reference's own check (LinkList/cMake/check.py): sort -> unsort gives ids in order
    ids = np.arange(r.shape[0], dtype=np.uint32)
    assert np.array_equal(ids[got["perm"]][got["inv_perm"]], ids)


@pytest.mark.parametrize("dims,n,hfac", [(3, 20, 2.0), (2, 90, 3.0), (3, 33, 1.3)])
def test_linklist_dam_break(oracle, dims, n, hfac):
    c = cases.dam_break(dims, n, hfac)
    ref = oracle.linklist(c["r"], dims, 2.0, c["h"])
    got = run_linklist(dims, c["r"], 2.0, c["h"])
    assert_ll_equal(ref, got)


def test_linklist_edge_cases(oracle):
    rng = np.random.default_rng(7)
    for N in (1, 2, 31, 4096, 4097):
        r = np.zeros((N, 4), np.float32)
        r[:, :3] = rng.uniform(-1, 1, (N, 3))
        if N > 2:
            r[1] = r[0]  # coincident
        ref = oracle.linklist(r, 3, 2.0, 0.05)
        got = run_linklist(3, r, 2.0, 0.05)
        assert_ll_equal(ref, got)


def test_linklist_ihoc_growth(ctx3):
    """LinkList::allocate (LinkList.cpp:234-271): ihoc is grown when n_cells.w rises."""
    N = 100
    r = np.zeros((N, 4), np.float32)
    r[:, :3] = np.random.default_rng(1).uniform(0, 1, (N, 3))
    d_r = ctx3.array(r)
    icell, perm, inv = (ctx3.empty(N, np.uint32) for _ in range(3))
    _, _, nc1, ihoc = ctx3.linklist(d_r, 2.0, 0.1, icell, None, perm, inv)
    r[0, :3] = 9.0
    d_r.set(r)
    _, _, nc2, ihoc2 = ctx3.linklist(d_r, 2.0, 0.1, icell, ihoc, perm, inv)
    assert nc2[3] > nc1[3] and ihoc2.shape[0] >= nc2[3]


@pytest.mark.parametrize("d", ["2D", "3D"])
def test_radix_sort_reference_inputs(golden, ctx3, d):
    """tests/{2D,3D}/RadixSort: keys in [0,1000) with ties; the reference's check.py
    asserts sortedness and perm / inverse-perm consistency; we also assert stability."""
    f_orig = golden[f"radixsort_{d}_f"]
    n = f_orig.size
    keys, perm, inv = ctx3.array(f_orig), ctx3.empty(n, np.uint32), ctx3.empty(n, np.uint32)
    ctx3.radix_sort(keys, 0, perm, inv)
    f, p, ip = keys.get(), perm.get(), inv.get()
    assert np.array_equal(f, np.sort(f_orig))
    assert not np.array_equal(f_orig, np.sort(f_orig))
    assert np.array_equal(f, f_orig[p])
    assert np.array_equal(f[ip], f_orig)
    assert np.array_equal(p, np.argsort(f_orig, kind="stable"))


# key_max picks the digits (linklist.cu make_plan): 1 pass of 8 bits; 2 x 8; 2 x 9 (the dam break's
# n_cells.w); 2 x 10; 2 x 11; 3 passes (11 + 11 + 10 for the full 32 bits, 8 + 8 + 8 for 23); sizes on
# and around the tile of 4096 keys, and two more tiles than one wave of CTAs holds
@pytest.mark.parametrize("n,key_max", [(1, 0), (2, 2), (33, 2), (255, 7), (4096, 256), (4097, 257),
                                       (8192, 1 << 18), (100003, 0), (1 << 20, 70000),
                                       (300001, 1 << 20), (123457, 1 << 22), (70001, 1 << 23),
                                       (3 * 1000 * 1000 + 17, 150000)])
def test_radix_sort_random(ctx3, n, key_max):
    rng = np.random.default_rng(n)
    hi = key_max if key_max else 2 ** 32
    k = rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32)
    keys, perm, inv = ctx3.array(k), ctx3.empty(n, np.uint32), ctx3.empty(n, np.uint32)
    ctx3.radix_sort(keys, key_max, perm, inv)
    ref = np.argsort(k, kind="stable").astype(np.uint32)
    assert np.array_equal(perm.get(), ref)
    assert np.array_equal(keys.get(), k[ref])
    ip = inv.get()
    assert np.array_equal(ip[ref], np.arange(n, dtype=np.uint32))
    # many equal keys (stability across tiles) and without the optional outputs; the context's
    # scratch is reused from the call above (digit totals / tickets cleaned by the library)
    k2 = (k % np.uint32(5)).astype(np.uint32) if key_max != 2 else k
    keys.set(k2)
    ctx3.radix_sort(keys, key_max, perm, None)
    ref2 = np.argsort(k2, kind="stable").astype(np.uint32)
    assert np.array_equal(perm.get(), ref2) and np.array_equal(keys.get(), k2[ref2])


def test_scatter_fields_and_fill(ctx3, oracle):
    rng = np.random.default_rng(3)
    N = 5000
    idx = rng.permutation(N).astype(np.uint32)
    a = rng.normal(size=(N, 4)).astype(np.float32)
    b = rng.integers(0, 100, N).astype(np.int32)
    m = rng.normal(size=(N, 16)).astype(np.float32)
    v2 = rng.normal(size=(N, 2)).astype(np.float32)
    d_idx = ctx3.array(idx)
    srcs = [ctx3.array(x) for x in (a, b, m, v2)]
    dsts = [ctx3.empty(x.shape, x.dtype) for x in (a, b, m, v2)]
    ctx3.scatter_fields(d_idx, list(zip(srcs, dsts)))
    for x, d in zip((a, b, m, v2), dsts):
        assert np.array_equal(d.get(), oracle.scatter(x, idx))
    pat = np.array([1.5, -2.0, 3.0, 0.0], np.float32)
    ctx3.fill(dsts[0], pat.tobytes())
    assert np.array_equal(dsts[0].get(), np.tile(pat, (N, 1)))


def test_reduce(ctx3):
    rng = np.random.default_rng(5)
    for n in (1, 1000, 1 << 20):
        x = rng.normal(size=n).astype(np.float32)
        d = ctx3.array(x)
        assert ctx3.reduce(_lib.OP_MIN, d) == x.min()
        assert ctx3.reduce(_lib.OP_MAX, d) == x.max()
        s = ctx3.reduce(_lib.OP_SUM, d)
        # the reference's own tolerance for sums is 1e-2 absolute on 500 values
        # (tests/3D/Reduction/cMake/check.py:23); scale it with sqrt(n)
        assert abs(s - x.astype(np.float64).sum()) <= 1e-5 * np.abs(x).sum() + 1e-6
        v = rng.normal(size=(n, 4)).astype(np.float32)
        dv = ctx3.array(v)
        assert np.array_equal(ctx3.reduce(_lib.OP_MIN, dv), v.min(0))
        assert np.array_equal(ctx3.reduce(_lib.OP_MAX, dv), v.max(0))
        u = rng.integers(0, 1000, n).astype(np.uint32)
        du = ctx3.array(u)
        assert ctx3.reduce(_lib.OP_MAX, du) == u.max()
        assert ctx3.reduce(_lib.OP_SUM, du) == u.sum(dtype=np.uint64) % 2 ** 32


@pytest.mark.parametrize("dims", [2, 3])
def test_linklist_matches_the_reference_tool_kernels(dims):
    """aqc_linklist_build against tests/golden/linklist_tool_outputs.npz: the cell of every particle and the
    head-of-cell table as the reference's OWN LinkList.cl.in kernels computed them (iCell, iHoc, linkList behind
    oracle/ref_shim, tests/golden/make_golden_linklist.py) for the reference's LinkList test particles, a dam break
    and random positions -- no oracle in between.  Bit-exact; the sorted order is the stable one."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_linklist as mk
    G = np.load(os.path.join(here, "golden", "linklist_tool_outputs.npz"))
    for name, r, h in mk.inputs(dims):
        key = "%dD_%s" % (dims, name)
        got = run_linklist(dims, r, 2.0, h)
        want = G[key + "_icell_unsorted"]
        assert np.array_equal(np.asarray(got["ncells"], np.uint32), G[key + "_ncells"]), key
        assert np.array_equal(got["icell"], want[got["perm"]]), key
        assert np.array_equal(got["perm"], np.argsort(want, kind="stable")), key
        assert np.array_equal(got["inv_perm"][got["perm"]], np.arange(r.shape[0], dtype=np.uint32)), key
        assert np.array_equal(got["ihoc"], G[key + "_ihoc"]), key
