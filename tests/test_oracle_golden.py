"""CPU: the oracle's link-list / radix sort against the reference's own test
inputs and property checks:

  tests/{2D,3D}/LinkList/cMake/check.py   -- saved ids are 0..N-1 in order, i.e.
                                             sort -> unsort is the identity
  tests/{2D,3D}/RadixSort/cMake/check.py  -- sortedness, f[i] == f_orig[perm[i]],
                                             f[inv[i]] == f_orig[i]
  tests/{2D,3D}/Reduction/cMake/check.py  -- sum within 1e-2, max within 1e-5

plus the definitions the reference's kernels rely on (LinkList.cl.in:32-113):
every particle is found through ihoc/icell of its own cell, cells are contiguous
runs, empty cells hold N, and the 27 (9) cell walk of BEGIN_NEIGHS finds every
particle closer than support*h (brute force)."""
import numpy as np
import pytest

import cases


def _check_structure(r, dims, h, ll):
    N = r.shape[0]
    icell, ihoc, perm, inv = ll["icell"], ll["ihoc"], ll["perm"], ll["inv_perm"]
    nc = ll["ncells"]
    # LinkList.cpp:185-232
    span = (ll["rmax"][:dims] - ll["rmin"][:dims]).astype(np.float32)
    n = (span / np.float32(np.float32(2.0) * np.float32(h))).astype(np.uint64) + 6
    assert np.array_equal(nc[:dims], n.astype(np.uint32))
    assert nc[3] == int(np.prod(n))
    # permutation pair (RadixSort.cl.in:295-296, 313-323) and stability
    assert np.array_equal(np.sort(perm), np.arange(N, dtype=np.uint32))
    assert np.array_equal(inv[perm], np.arange(N, dtype=np.uint32))
    # icell (LinkList.cl.in:54-85) of the ORIGINAL order, recomputed in numpy
    idist = np.float32(1.0) / (np.float32(2.0) * np.float32(h))
    c = ((r[:, :dims] - ll["rmin"][:dims]).astype(np.float32) * idist).astype(np.float32)
    c = c.astype(np.uint32) + np.uint32(3)
    cell = c[:, 0] - 1 + (c[:, 1] - 1) * nc[0]
    if dims == 3:
        cell = cell + (c[:, 2] - 1) * nc[0] * nc[1]
    cell = cell.astype(np.uint32)
    assert np.array_equal(icell, cell[perm])
    assert np.array_equal(perm, np.argsort(cell, kind="stable").astype(np.uint32))
    assert np.all(np.diff(icell.astype(np.int64)) >= 0)
    # heads (LinkList.cl.in:32-42, 92-113)
    first = np.full(nc[3], N, np.uint32)
    heads = np.flatnonzero(np.concatenate([[True], icell[1:] != icell[:-1]]))
    if N > 1:   # linkList runs on N-1 work-items (LinkList.cl.in:92-113): N == 1 sets no head
        first[icell[heads]] = heads
    assert np.array_equal(ihoc[:nc[3]], first)


def _brute_neighbours(r, dims, h, ll):
    """BEGIN_NEIGHS (types/3D.h:197-219) must visit every j with |r_ij| < support*h."""
    N = r.shape[0]
    icell, ihoc, nc, perm = ll["icell"], ll["ihoc"], ll["ncells"], ll["perm"]
    rs = r[perm][:, :dims].astype(np.float64)
    cut = 2.0 * h
    for i in range(0, N, max(1, N // 60)):
        d = np.sqrt(((rs - rs[i]) ** 2).sum(1))
        want = set(np.flatnonzero(d < cut * (1 - 1e-6)).tolist())
        got = set()
        kz = (-1, 0, 1) if dims == 3 else (0,)
        for ci in (-1, 0, 1):
            for cj in (-1, 0, 1):
                for ck in kz:
                    cc = int(icell[i]) + ci + cj * int(nc[0]) + ck * int(nc[0]) * int(nc[1])
                    j = int(ihoc[cc])
                    while j < N and icell[j] == cc:
                        got.add(j)
                        j += 1
        assert want <= got


@pytest.mark.parametrize("d", ["2D", "3D"])
def test_linklist_reference_inputs(golden, oracle, d):
    dims = int(d[0])
    r1 = golden[f"linklist_{d}_r"]
    # tests/3D/LinkList/cMake/main.xml loads the same file in two sets, h = 0.1
    r = np.ascontiguousarray(np.concatenate([r1, r1]))
    ll = oracle.linklist(r, dims, 2.0, 0.1)
    _check_structure(r, dims, 0.1, ll)
    _brute_neighbours(r, dims, 0.1, ll)
    # the reference's check.py: ids saved after sort -> unsort come back in order
    ids = np.arange(r.shape[0], dtype=np.uint32)
    sorted_ids = oracle.scatter(ids, ll["inv_perm"])        # basic/Sort.cl
    assert np.array_equal(oracle.scatter(sorted_ids, sorted_ids), ids)  # UnSort by id
    # coincident particles (Create.py:41-47) share a cell and keep their input order
    same = np.flatnonzero(np.all(r1[1:] == r1[:-1], axis=1))
    assert same.size > 0
    for k in same:
        assert ll["inv_perm"][k + 1] == ll["inv_perm"][k] + 1


@pytest.mark.parametrize("d", ["2D", "3D"])
def test_radix_sort_reference_inputs(golden, oracle, d):
    f_orig = golden[f"radixsort_{d}_f"]
    n = f_orig.size
    keys = f_orig.copy()
    perm, inv = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    oracle.call("radix_sort", keys, n, perm, inv)
    assert np.all(keys[1:] >= keys[:-1])                      # check.py:14-15
    assert np.array_equal(keys, f_orig[perm])                 # check.py:16-17
    assert np.array_equal(keys[inv], f_orig)                  # check.py:18-19
    assert np.array_equal(perm, np.argsort(f_orig, kind="stable"))


@pytest.mark.parametrize("dims,n,hfac", [(3, 12, 2.0), (2, 50, 3.0), (3, 9, 1.3), (2, 30, 4.0)])
def test_linklist_dam_break(oracle, dims, n, hfac):
    c = cases.dam_break(dims, n, hfac)
    ll = oracle.linklist(c["r"], dims, 2.0, c["h"])
    _check_structure(c["r"], dims, c["h"], ll)
    _brute_neighbours(c["r"], dims, c["h"], ll)


def test_linklist_edge_cases(oracle):
    rng = np.random.default_rng(7)
    for N in (1, 2, 31, 1025):
        r = np.zeros((N, 4), np.float32)
        r[:, :3] = rng.uniform(-1, 1, (N, 3))
        if N > 2:
            r[1] = r[0]
        _check_structure(r, 3, 0.05, oracle.linklist(r, 3, 2.0, 0.05))
    with pytest.raises(RuntimeError):
        oracle.linklist(np.zeros((3, 4), np.float32), 3, 2.0, 0.0)   # LinkList.cpp:193-198


def test_reduction_reference_tolerances(golden, oracle):
    """tests/3D/Reduction: sum tol 1e-2 absolute, max tol 1e-5 on 500 values."""
    import ctypes as C
    x = np.ascontiguousarray(golden["radixsort_3D_r"][:, 0])
    L = oracle.lib()
    for wg in (64, 256, 1024):
        s = L.aqo_reduce_sum_tree(oracle._arg(x), C.c_uint32(x.size), C.c_uint32(wg))
        assert abs(s - x.astype(np.float64).sum()) < 1e-2
    assert abs(L.aqo_reduce_max(oracle._arg(x), C.c_uint32(x.size)) - x.max()) < 1e-5
    assert L.aqo_reduce_min(oracle._arg(x), C.c_uint32(x.size)) == x.min()


def test_define_rounding(oracle):
    """CalcServer.cpp:245-257: evaluated defines are printed with %#G (6 digits)."""
    L = oracle.lib()
    assert L.aqo_define_round6(np.float32(0.0123456789)) == np.float32(0.0123457)
    assert L.aqo_define_round6(np.float32(1234567.0)) == np.float32(1.23457e6)
    d = oracle.make_defs(3, 0.03)
    assert d.H == np.float32(0.03) and d.SUPPORT == 2.0
    assert d.CONW == np.float32(float("%#G" % (np.float32(1.0) / np.float32(0.03) ** 3)))


# ---- value-level goldens of the ideal-gas family (tests/golden/ideal_gas_outputs.npz: outputs of the REFERENCE's
# own scripts on seeded inputs, generated by tests/golden/make_golden_ideal_gas.py in the build container) ----
def _ideal_gas_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ideal_gas_outputs.npz"))


@pytest.mark.parametrize("dims", [2, 3])
def test_ideal_gas_restatements_equal_the_reference_outputs(oracle, dims):
    """The C restatements of resources/Scripts/cfd/ideal_gas (element-wise kernels, every time scheme, the
    symmetry copy) against what the reference's scripts produced: the same bits, without the reference tree."""
    import oracle.oracle as O
    from test_oracle_vs_reference import _ideal_gas_state
    G = _ideal_gas_golden()
    _, b, N = _ideal_gas_state(dims, 13)
    D = O.make_defs(dims, b["h"])
    c = oracle.call
    c("ig_eos", b["iset"], b["imove"], b["rho"], b["eint"], b["p"], b["gamma"], N)
    c("ig_rates", b["imove"], b["rho"], b["p"], b["div_u"], b["deintdt"], N)
    c("ig_timestep", D, b["dt_var"], b["imove"], b["iset"], b["u"], b["rho"], b["p"], N, b["dt"], b["dt_min"],
      b["courant"], b["div_u"], b["grad_p"], b["gamma"])
    c("ig_mp_predictor", b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N)
    c("ig_riemann_rates", b["imove"], b["work_density"], b["deintdt"], N)
    c("ig_mp_midpoint", b["imove"], b["eint_in"], b["deintdt"], b["eint"], N, b["dt"])
    c("ig_mp_relax", b["imove"], b["deintdt_in"], b["deintdt"], N, b["relax_midpoint"])
    c("ig_mp_corrector", b["imove"], b["eint_in"], b["deintdt"], b["eint"], N, b["dt"])
    b["eint_in"][...] = b["eint"]
    c("ig_sort", b["eint_in"], b["eint"], b["deintdt"], b["deintdt_in"], b["id_sorted"], N)
    c("ig_mp_predictor", b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N)
    c("ig_euler_corrector", b["imove"], b["eint"], b["deintdt"], N, b["dt"])
    b["deintdt"][...] = b["work_density"]
    c("ig_ie_corrector", b["imove"], b["deintdt"], b["deintdt_in"], b["eint"], N, b["dt"])
    c("ig_ie_predictor", b["imove"], b["eint"], b["deintdt"], b["eint_in"], b["deintdt_in"], N, b["dt"])
    c("ig_sym_set", b["mirror_src"], b["eint_in"], b["deintdt_in"], b["deintdt"], N)
    for k in ("p", "deintdt", "dt_var", "eint", "eint_in", "deintdt_in"):
        assert G["elementwise_%dD_%s" % (dims, k)].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize("dims,n,hfac", [(2, 40, 3.0), (3, 10, 2.0)])
def test_riemann_restatement_equals_the_reference_outputs(oracle, dims, n, hfac):
    import oracle.oracle as O
    import pipeline
    from test_oracle_vs_reference import riemann_inputs
    G = _ideal_gas_golden()
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    x = riemann_inputs(s)
    oracle.call("ig_riemann_interactions", O.make_defs(dims, s["h"]), pipeline._ll(s), x["iset"], s["imove"], s["r"],
                x["u"], s["rho"], s["m"], x["p"], x["grad_p"], x["div_u"], x["work_density"], x["gamma"])
    for k in ("grad_p", "div_u", "work_density"):
        assert G["riemann_%dD_%s" % (dims, k)].tobytes() == x[k].tobytes(), k


@pytest.mark.parametrize("dims,n,hfac", [(3, 10, 3.0), (2, 40, 4.0)])
def test_hot_path_restatement_equals_the_reference_outputs(oracle, dims, n, hfac):
    """tests/golden/hotpath_sweeps_outputs.npz: what the reference's own scripts (behind oracle/ref_shim, build
    container, tests/golden/make_golden_hotpath.py) produced for the hot-path sequence of tests/pipeline.py --
    EOS, MLS, Shepard, Interactions, sensors, delta-SPH, the BIe boundary terms, Rates, residuals, TimeStep -- on
    the cell-sorted dam-break state: the oracle gives the same bits, output by output, with neither the reference
    tree nor oracle/_ref at hand (what tests/test_oracle_vs_reference.py checks live where they are)."""
    import os
    import pipeline
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_sweeps_outputs.npz"))
    case = cases.dam_break(dims, n, hfac)
    s = pipeline.oracle_linklist_and_sort(case)
    got = pipeline.oracle_sweeps(s)
    keys = [k[len("sweeps_%dD_" % dims):] for k in G.files if k.startswith("sweeps_%dD_" % dims)]
    assert len(keys) >= 30 and set(keys) == set(got)
    bad = [k for k in keys if not np.array_equal(G["sweeps_%dD_%s" % (dims, k)], np.asarray(got[k]), equal_nan=True)]
    assert not bad, bad
    assert np.abs(got["grad_p"]).max() > 0 and np.abs(got["lap_p"]).max() > 0


@pytest.mark.parametrize("dims", [2, 3])
def test_linklist_restatement_equals_the_reference_tool_kernels(oracle, dims):
    """tests/golden/linklist_tool_outputs.npz: icell of every particle and the head-of-cell table as the reference's
    own LinkList.cl.in kernels computed them (behind oracle/ref_shim, tests/golden/make_golden_linklist.py) for the
    reference's LinkList test particles, a dam break and random positions.  The oracle's link-list (cell hash, stable
    sort, heads) gives the same integers -- bit-exact, with neither the reference tree nor oracle/_ref at hand."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_linklist as mk
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linklist_tool_outputs.npz"))
    for name, r, h in mk.inputs(dims):
        key = "%dD_%s" % (dims, name)
        ll = oracle.linklist(r, dims, 2.0, h)
        assert np.array_equal(ll["ncells"], G[key + "_ncells"]), key
        want = G[key + "_icell_unsorted"]
        assert np.array_equal(ll["icell"], want[ll["perm"]]), key                    # the same cells ...
        assert np.array_equal(ll["perm"], np.argsort(want, kind="stable")), key    # ... in the stable order
        ncw = int(ll["ncells"][3])
        assert np.array_equal(ll["ihoc"][:ncw], G[key + "_ihoc"]), key
        assert (G[key + "_ihoc"] < r.shape[0]).sum() == len(np.unique(want)), key


@pytest.mark.parametrize("dims", [2, 3])
def test_time_scheme_restatements_equal_the_reference_outputs(oracle, dims):
    """tests/golden/time_scheme_outputs.npz: the reference's basic/time_scheme/{midpoint, euler, improved_euler}.cl
    and basic/Domain.cl (removal branch included: NaN and out-of-box positions) run one after the other on a seeded
    state (tests/golden/make_golden_elementwise.py); the restatements, run the same way, leave the same bits."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_elementwise as mk
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "time_scheme_outputs.npz"))
    v, N = mk.state(dims)
    lit = {"N": N, "dt": mk.DT, "relax": mk.RELAX, "dims": dims}
    for script, entry, outs, fn, args in mk.SEQUENCE:
        if script is None:
            mk.spoil(v)
            continue
        oracle.call(fn, *[lit[a] if a in lit else v[a] for a in args])
        for k in outs:
            want = G["%dD_%s_%s_%s" % (dims, os.path.basename(script)[:-3], entry, k)]
            assert np.array_equal(want, v[k], equal_nan=True), (script, entry, k)
    assert (v["imove"] == -256).any()        # Domain did remove particles


@pytest.mark.parametrize("dims", [2, 3])
def test_open_boundary_kernels_match_the_reference_outputs(oracle, dims):
    """oracle/aqo_kernels.c: aqo_inlet_* / aqo_outlet_* / aqo_portal_* against the committed outputs of the
    reference's own scripts (cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl, Portal/Mirror.cl run behind
    oracle/ref_shim by tests/golden/make_golden_open_boundary.py): the same bits after every kernel, without
    the reference tree."""
    import os
    import open_boundary_common as ob
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "open_boundary_outputs.npz"))
    case, v = ob.state(dims)
    D = oracle.make_defs(dims, case["h"])
    b = ob.args_of(v)
    for step, key in enumerate(ob.STEPS):
        ob.oracle_step(oracle, D, dims, key, b)
        for k in ob.WRITES[key]:
            assert b[k].tobytes() == G["%dD_step%d_%s" % (dims, step, k)].tobytes(), (key, k)
    ob.checks(v, b, dims)
