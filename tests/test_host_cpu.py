"""CPU (no GPU): the C-ABI libraries load and export every symbol the headers
declare, they refuse to run without a device (no CPU fallback), and the C++
host's XML front-end / Variables / Tokenizer behave like the reference's
(State.cpp:276-389,739-1217; Variable.cpp:1321-1545; tests/*/SetScalar)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

from aquagpusph_b200 import _lib, cases, casegen, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aq[ch]_\w+)\s*\(", txt)))


def test_aquacuda_exports_every_declared_symbol():
    names = _declared("aquacuda.h")
    assert len(names) >= 30
    L = ctypes.CDLL(os.path.join(ROOT, "aquagpusph_b200", "libaquacuda.so"))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_aquahost_exports_every_declared_symbol():
    names = _declared("aquahost.h")
    L = host.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(host.SYMBOLS) == set(names)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it never routes to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.AquaError, match="no CPU fallback"):
        _lib.Context(0, dims=3, h=0.1)
    c = cases.spheric2_dam_break(2000, 3.0)
    with pytest.raises(host.HostError, match="no CPU fallback"):
        casegen.load("spheric2_dambreak_3d", c, (c["N"] - 8, 8))
    for mod in ("_lib", "host", "casegen", "cases", "build"):
        src = open(os.path.join(ROOT, "aquagpusph_b200", mod + ".py")).read()
        assert "oracle" not in src.replace("the oracle", ""), mod + ".py must not touch oracle/"


def test_kernel_registry_matches_reference_signatures():
    """Appendix D of SURVEY.md: argument names/order the Kernel tool binds by name
    (Kernel.cpp:497-556).  A few representative hot-path entries."""
    L = _lib.lib()
    want = {
        ("cfd/Interactions.cl", "entry"): ["imove", "r", "u", "rho", "m", "p", "grad_p", "lap_u",
                                           "div_u", "N", "icell", "ihoc", "n_cells"],
        ("cfd/Rates.cl", "entry"): ["iset", "imove", "rho", "grad_p", "lap_u", "div_u", "dudt",
                                    "drhodt", "visc_dyn", "N", "g"],
        ("cfd/TimeStep.cl", "entry"): ["imove", "u", "dt_var", "N", "dt", "dt_min", "courant",
                                       "dt_Ma", "h"],
        ("basic/time_scheme/midpoint.cl", "corrector"): ["imove", "r_in", "r", "u_in", "u", "dudt",
                                                         "rho_in", "rho", "drhodt", "N", "dt"],
        ("basic/Sort.cl", "stage2"): ["rho_in", "rho", "m_in", "m", "u_in", "u", "dudt", "dudt_in",
                                      "drhodt", "drhodt_in", "id_sorted", "N"],
        ("basic/EOS.cl", "entry"): ["iset", "imove", "rho", "p", "refd", "N", "cs", "p0"],
    }
    for (script, entry), names in want.items():
        # presets write "../Scripts/<path>" or an absolute resources path
        kid = L.aqc_kernel_lookup(("/x/resources/Scripts/" + script).encode(), entry.encode(), 3)
        assert kid >= 0, (script, entry)
        n = L.aqc_kernel_nargs(kid)
        args = L.aqc_kernel_args(kid)
        assert [args[k].name.decode() for k in range(n)] == names
    assert L.aqc_kernel_lookup(b"cfd/NoSuch.cl", b"entry", 3) == -3


# ---- Variables / Tokenizer -----------------------------------------------------
def test_set_scalar_sequence_of_the_reference_test():
    """tests/3D/SetScalar/cMake/main.xml + check.py: h=1, i=-1, v=(3,4,5,6);
    set float, recursive float, vector swap, float*int, vector*float ->
    'i h v' == '-1 -4 (-16,-0.75,5,6)'."""
    d = "float h=1.0;int i=-1;vec v=3.0, 4.0, 5.0, 6.0"
    h = host.evaluate("2.0", decls=d)
    h = host.evaluate("2.0 * h", decls=d.replace("h=1.0", "h=%r" % float(h)))
    assert h == 4.0
    v = host.evaluate("v_y,v_x,v_z,v_w", "vec", d, n=4)
    assert list(v) == [4.0, 3.0, 5.0, 6.0]
    d2 = "float h=4.0;int i=-1;vec v=4.0, 3.0, 5.0, 6.0"
    h = host.evaluate("h * i", decls=d2)
    assert h == -4.0
    v = host.evaluate("v_x * h,v_y / h,v_z,v_w", "vec", d2.replace("h=4.0", "h=-4.0"), n=4)
    assert list(v) == [-16.0, -0.75, 5.0, 6.0]


def test_expression_semantics():
    ev = host.evaluate
    assert ev("1/3") == np.float32(1.0 / 3.0)              # double, narrowed once
    assert ev("2^3^2") == 512.0                           # right associative
    assert ev("-2^2") == -4.0
    assert ev("a < b ? a : b", decls="float a=3;float b=2") == 2.0
    assert ev("(a > 1) && !(b > 5) || 0", decls="float a=3;float b=2") == 1.0
    assert ev("sqrt(16) + abs(-1) + max(1,2,3) + min(4,2)") == 10.0
    assert ev("2.f * 0.5f") == 1.0                        # OpenCL-style literals
    assert ev("pi") == np.float32(np.pi)
    assert ev("7 % 4") == 3.0
    assert ev("7.9", "unsigned int", dtype=np.uint32) == 7  # truncation like numeric_cast
    assert ev("0.25*h/cs", decls="float dr=0.01;float hfac=3;float h=hfac*dr;float cs=40") == \
        np.float32(0.25 * float(np.float32(3 * 0.01)) / 40)
    nc = ev("n_x*n_y, 2, 3, n_w", "uivec4", "uivec4 n=4,5,6,120", dtype=np.uint32, n=4)
    assert list(nc) == [20, 2, 3, 120]
    # vec is 2 floats in 2-D and 4 in 3-D (Variable.cpp:1547-1624)
    assert ev("1,2", "vec", dims=2, n=2).tolist() == [1.0, 2.0]
    with pytest.raises(host.HostError, match="Invalid number of fields"):
        ev("1,2", "vec", dims=3, n=4)
    with pytest.raises(host.HostError, match="cannot be found"):
        ev("2*nope")
    with pytest.raises(host.HostError):
        ev("2*(3")
    with pytest.raises(host.HostError, match="overflows"):
        ev("-1", "unsigned int", dtype=np.uint32)


# ---- XML front-end ---------------------------------------------------------------
KEY_ORDER_3D = ["predictor", "Domain", "link-list", "sort stage1", "sort stage2", "EOS",
                "midpoint loop", "cfd Shepard", "cfd interactions", "cfd lap p", "cfd rates",
                "midpoint residual", "midpoint loop end", "corrector", "cfd variable time step",
                "cfd minimum time step", "cfd check time step", "t = t + dt", "iter += 1"]


def _parse(txt, dims):
    d = tempfile.mkdtemp(prefix="aqua_xml_")
    p = os.path.join(d, "case.xml")
    open(p, "w").write(txt)
    return host.Simulation(p, dims=dims, parse_only=True)


def test_resolved_dam_break_3d_pipeline():
    c = cases.spheric2_dam_break(2000, 3.0)
    sim = _parse(casegen.instantiate("spheric2_dambreak_3d", c, (c["N"] - 8, 8),
                                     keep_reports=True), 3)
    tools = sim.tools()
    names = [t[0] for t in tools]
    types = [t[1] for t in tools]
    # 116 tools (indices 0..115) followed by the example's 6 reports
    assert len(tools) == 122 and names[115] == "End" and types[116:] == [
        "report_screen", "report_file", "report_performance", "report_particles", "report_screen",
        "report_file"]
    assert types.count("link-list") == 1 and types.count("while") == 1 and types.count("end") == 2
    assert (names[4], names[44], names[50], names[70], names[91], names[93], names[98]) == (
        "link-list", "midpoint loop", "cfd interactions", "cfd rates", "midpoint loop end",
        "corrector", "cfd minimum time step")
    pos = []
    for key in KEY_ORDER_3D:
        hits = [i for i, n in enumerate(names) if n.lower() == key.lower()]
        assert hits, "tool %r missing; have %r" % (key, names)
        pos.append(hits[0])
    assert pos == sorted(pos), list(zip(KEY_ORDER_3D, pos))
    # the midpoint sub-iteration encloses the neighbour sweeps
    w, e = types.index("while"), names.index("midpoint loop end")
    inner = names[w:e]
    assert "cfd interactions" in inner and "cfd rates" in inner
    sweeps = [n for n, t in tools if t == "kernel"]
    assert len(sweeps) == 37
    # flat XML round trip (State::write -> State::parse) keeps the pipeline
    out = os.path.join(tempfile.mkdtemp(), "resolved.xml")
    sim.write_resolved(out)
    again = host.Simulation(out, dims=3, parse_only=True).tools()
    assert again == tools


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"),
                    reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("name,src,dims", [
    ("spheric2_dambreak_3d", "examples/3D/spheric_testcase2_dambreak/src/templates", 3),
    ("spheric5_dambreak_2d", "examples/2D/spheric_testcase5_dambreak/src/templates", 2),
    ("spheric9_tld_2d", "examples/2D/spheric_testcase9_tld/src/templates", 2),
    ("spheric3_liddriven_2d", "examples/2D/spheric_testcase3_liddriven/src/templates", 2),
    ("souto2012_standingwave_2d", "examples/2D/souto_etal_2012_standingwave/src/templates", 2),
    ("shock_point_2d", "examples/2D/shock_point/src/templates", 2),
])
def test_committed_templates_match_the_reference_examples(name, src, dims):
    """The committed resolved templates are what our front-end makes of the
    reference's unchanged Main.xml + presets (include/prefix/insert semantics of
    State.cpp:739-1217, lenient handling of the malformed kernel presets)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("resolve_case",
                                                  os.path.join(ROOT, "tools", "resolve_case.py"))
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    out = tempfile.mkdtemp()
    old = rc.OUT
    rc.OUT = out
    try:
        rc.resolve(name, src, dims)
    finally:
        rc.OUT = old
    fresh = open(os.path.join(out, name + ".xml")).read()
    committed = open(os.path.join(ROOT, "aquagpusph_b200", "cases_xml", name + ".xml")).read()

    def tools(t):
        return re.findall(r'<Tool [^>]*name="([^"]*)"[^>]*type="([^"]*)"', t)
    assert tools(fresh) == tools(committed)
    n = len(tools(committed))
    # 116 tools + the example's reports (3-D), 57 (2-D)
    assert (dims == 3 and n >= 116) or (dims == 2 and n >= 55)


def test_tool_placement_semantics():
    """<Tool action=insert before/after=..., remove, replace> and ifdef gating
    (State.cpp:778-1021) on a hand-written pipeline."""
    xml = """<?xml version="1.0" ?>
<sphInput>
  <Variables>
    <Variable name="h" type="float" value="0.1" />
    <Variable name="a" type="float" value="1" />
  </Variables>
  <Definitions>
    <Define name="WITH_X" value="1" />
  </Definitions>
  <Tools>
    <Tool action="add" name="t1" type="set_scalar" in="a" value="1" />
    <Tool action="add" name="t3" type="set_scalar" in="a" value="3" />
    <Tool action="insert" before="t3" name="t2" type="set_scalar" in="a" value="2" />
    <Tool action="insert" after="t3" name="t4" type="set_scalar" in="a" value="4" />
    <Tool action="add" name="gone" type="set_scalar" in="a" value="0" />
    <Tool action="remove" name="gone" type="dummy" />
    <Tool action="replace" name="t1" type="set_scalar" in="a" value="10" />
    <Tool action="add" name="gated in" type="dummy" ifdef="WITH_X" />
    <Tool action="add" name="gated out" type="dummy" ifdef="WITHOUT_X" />
    <Tool action="try_remove" name="never existed" type="dummy" />
  </Tools>
  <Timing><Option name="End" type="Steps" value="1" /></Timing>
  <ParticlesSet n="4" />
</sphInput>
"""
    names = [t[0] for t in _parse(xml, 2).tools()]
    assert names == ["t1", "t2", "t3", "t4", "gated in"]
    with pytest.raises(host.HostError):
        _parse(xml.replace('before="t3"', 'before="nowhere"'), 2)


def test_whole_kernel_registry_against_the_reference_scripts():
    """Every registered kernel against the signature the reference's Kernel tool would reflect from
    its .cl script with clGetKernelArgInfo (Kernel.cpp:497-556; fixture: tests/golden/
    kernel_signatures.json, made by make_kernel_signatures.py from /root/reference): same argument
    names in the same order, pointers where the script has pointers, AQC_ARG_ARRAY_IN exactly for the
    `const` / `__constant` ones and ARRAY_OUT / ARRAY_RO for the others."""
    import json
    sig = json.load(open(os.path.join(ROOT, "tests", "golden", "kernel_signatures.json")))
    L = _lib.lib()
    checked, own = 0, []
    for kid in range(L.aqc_kernel_count()):
        name = L.aqc_kernel_name(kid).decode()
        script, entry = name.split("::")
        key = script if script in sig else "examples:" + script
        if key not in sig or entry not in sig[key]:
            own.append(name)
            continue
        ref = sig[key][entry]
        args = L.aqc_kernel_args(kid)
        got = [(args[k].name.decode(), args[k].kind) for k in range(L.aqc_kernel_nargs(kid))]
        assert [g[0] for g in got] == [r[0] for r in ref], name
        for (gname, kind), (_, ptr, const) in zip(got, ref):
            assert (kind != _lib.ARG_SCALAR) == ptr, (name, gname)
            if ptr:
                assert (kind == _lib.ARG_ARRAY_IN) == const, (name, gname, kind)
        checked += 1
    assert checked >= 55
    # the only kernels without a reference script are this library's own: the pair-count diagnostic
    # and the remote delta-SPH / MLS terms of the slab pipelines (the reference's MPI preset has none)
    assert sorted(own) == ["aqua/MPIdeltaSPH.cl::copy_g", "aqua/MPIdeltaSPH.cl::full_lapp",
                           "aqua/MPIdeltaSPH.cl::lapp_corr", "aqua/MPIdeltaSPH.cl::mls",
                           "aqua/MPIdeltaSPH.cl::sort_g", "aqua/diag.cl::count_pairs"], own


# ---- device-side loops: the scalar programs (include/aquasvm.h, host/devloop.cpp) ----------------
def _loop_expressions():
    """Every set_scalar value, while / if / assert condition of the committed pipelines."""
    import glob
    import xml.etree.ElementTree as ET
    out = set()
    for f in glob.glob(os.path.join(ROOT, "aquagpusph_b200", "cases_xml", "*.xml")):
        for t in ET.parse(f).getroot().iter("Tool"):
            if t.get("type") == "set_scalar":
                out.add(t.get("value"))
            elif t.get("type") in ("while", "if", "assert"):
                out.add(t.get("condition"))
    return sorted(out)


def test_svm_programs_equal_the_host_evaluator():
    """SvmCompiler + aqs_run (what a recorded `while` runs on the device) give
    the bits of Variables::solve (SetScalar.cpp:146-195 through Tokenizer::solve): same grammar,
    double arithmetic one operation at a time, one narrowing per store."""
    ev, sv = host.evaluate, host.evaluate_svm
    d = ("float a=3;float b=2;float h=0.0123;float cs=45.5;int i=-7;unsigned int u=4000000000;"
         "vec v=3.0, 4.0, 5.0, 6.0;float relax=0.75;float r0=1.25e-3;float r1=3.5e-3;float z=0")
    exprs = ["1/3", "2^3^2", "-2^2", "a < b ? a : b", "(a > 1) && !(b > 5) || 0",
             "sqrt(16) + abs(-1) + max(1,2,3) + min(4,2)", "2.f * 0.5f", "pi", "7 % 4", "e^a",
             "0.25*h/cs", "i*h", "u/3", "v_x*v_y - v_z/v_w", "a - b - 1", "a / b / 3", "a - -b",
             "!a", "!z + !(a != 3)", "a <= 3 == 1", "r1 / r0 > 0.9 ? 0.5 + (1.0 - 0.5) * relax : relax",
             "i == 0 ? 1.0 : relax", "r1 < 1e-2 ? 30 : i", "pow(a, b) + log(a) + ln(a) + log2(a) + log10(a)",
             "floor(-1.5) + ceil(1.2) + round(2.5) + rint(3.5) + sign(-a) + sign(z)",
             "sin(a) + cos(b) + tan(h) + asin(0.5) + acos(0.5) + atan(a) + atan2(a, b)",
             "sinh(h) + cosh(h) + tanh(a) + exp(-b)", "avg(a, b, 4) + sum(a, b)", "a > b > 0",
             "a*b+h*cs", "h*cs+a*b", "(a+b)*(a-b)/(h+cs)", "1.e-8 + 2.E+3f", "+a", "-(-a)", "a != b != 0"]
    for x in exprs + _loop_expressions():
        names = set(re.findall(r"[A-Za-z_]\w*", x))
        dd = d
        for n in sorted(names - set("a b h cs i u v relax r0 r1 z pi e".split())):
            if re.search(r"\b%s\s*\(" % n, x) or n[:-2] == "v" or n in ("f", "E"):
                continue
            dd += ";float %s=%r" % (n, 0.3 + 0.001 * (sum(map(ord, n)) % 977))
        for ty, dt in (("float", np.float32), ("int", np.int32)):
            try:
                want = ev(x, ty, dd, dtype=dt)
            except host.HostError:
                with pytest.raises(host.HostError):
                    sv(x, ty, dd, dtype=dt)
                continue
            got = sv(x, ty, dd, dtype=dt)
            assert want.tobytes() == got.tobytes() or (np.isnan(want) and np.isnan(got)), (x, ty, want, got)
    # vector values: every component sees the OLD variable (Variables::solve stores at the end)
    for x in ("v_y,v_x,v_z,v_w", "v_x * h,v_y / h,v_z,v_w"):
        assert ev(x, "vec", d, n=4).tobytes() == sv(x, "vec", d, n=4).tobytes()
    assert sv("1,2", "vec", dims=2, n=2).tolist() == [1.0, 2.0]
    assert sv("7.9", "unsigned int", dtype=np.uint32) == 7
    nc = sv("n_x*n_y, 2, 3, n_w", "uivec4", "uivec4 n=4,5,6,120", dtype=np.uint32, n=4)
    assert list(nc) == [20, 2, 3, 120]
    # the same errors
    with pytest.raises(host.HostError, match="Invalid number of fields"):
        sv("1,2", "vec", dims=3, n=4)
    with pytest.raises(host.HostError, match="cannot be found"):
        sv("2*nope")
    with pytest.raises(host.HostError):
        sv("2*(3")
    with pytest.raises(host.HostError, match="overflows"):
        sv("-1", "unsigned int", dtype=np.uint32)
    with pytest.raises(host.HostError, match="overflows"):
        sv("3e9", "int", dtype=np.int32)
    with pytest.raises(host.HostError, match="expects 2 arguments"):
        sv("atan2(1)")


def _rand_expr(rng, names, depth=0):
    """A random expression of the tokenizer's grammar (host/tokenizer.hpp)."""
    r = rng.random()
    if depth > 4 or r < 0.22:
        if rng.random() < 0.5:
            return str(rng.choice(names))
        v = rng.choice([0, 1, 2, 3, 0.5, 0.25, 1.5, 7, 10, 1e-3, 2.5e2, 0.1, 3.7])
        return rng.choice(["%r" % float(v), "%g" % v, "%gf" % v if "e" not in "%g" % v else "%r" % float(v)])
    e = lambda: _rand_expr(rng, names, depth + 1)   # noqa: E731
    if r < 0.55:
        return "%s %s %s" % (e(), rng.choice(["+", "-", "*", "/", "+", "*"]), e())
    if r < 0.62:
        return "(%s)" % e()
    if r < 0.68:
        return "-%s" % e()
    if r < 0.72:
        return "!(%s)" % e()
    if r < 0.80:
        return "%s %s %s" % (e(), rng.choice(["<", ">", "<=", ">=", "==", "!="]), e())
    if r < 0.85:
        return "(%s) %s (%s)" % (e(), rng.choice(["&&", "||"]), e())
    if r < 0.92:
        return "(%s) ? %s : %s" % (e(), e(), e())
    f = rng.choice(["sqrt(abs(%s))", "abs(%s)", "min(%s, %s)", "max(%s, %s, %s)", "floor(%s)", "ceil(%s)",
                    "(%s)^2", "sign(%s)", "avg(%s, %s)", "7 %% (1 + abs(%s))"])
    return f % tuple(e() for _ in range(f.count("%s")))


def test_svm_programs_equal_the_host_evaluator_on_random_expressions():
    """400 random expressions of the tokenizer's grammar over float / int / unsigned / vector-component
    variables: SvmCompiler + aqs_run and Variables::solve give the same bits -- or fail together (a value
    that overflows the target type)."""
    rng = np.random.default_rng(2026)
    d = ("float a=3;float b=-2.5;float c=0.125;int i=-7;int k=12;unsigned int u=4000000000;unsigned int n=17;"
         "vec v=3.0, -4.0, 5.5, 6.0;float z=0;float big=3e9")
    names = ["a", "b", "c", "i", "k", "u", "n", "v_x", "v_y", "v_z", "v_w", "z", "big"]
    ran = failed = 0
    for _ in range(400):
        x = _rand_expr(rng, names)
        for ty, dt in (("float", np.float32), ("int", np.int32), ("unsigned int", np.uint32)):
            try:
                want = host.evaluate(x, ty, d, dtype=dt)
            except host.HostError:
                with pytest.raises(host.HostError):
                    host.evaluate_svm(x, ty, d, dtype=dt)
                failed += 1
                continue
            got = host.evaluate_svm(x, ty, d, dtype=dt)
            assert want.tobytes() == got.tobytes() or (np.isnan(want) and np.isnan(got)), (x, ty, want, got)
            ran += 1
    assert ran > 600 and failed > 20, (ran, failed)


def test_lane_schedule_orders_every_conflicting_pair():
    """The two-lane schedule of the device loops (host/devloop.cpp::scheduleLanes through
    aqh_lane_schedule), on random dependency sets: whatever lanes it picks, every pair of tools of which one
    writes what the other reads or writes (in rows of a common particle class) is ordered -- by the in-order
    lane they share, or by a chain of recorded events and waits -- in pipeline order; tools flagged for lane
    0 stay there; the last tool of lane 1 carries the event the pass joins on.  And it does use the second
    lane: two independent chains of sweeps end up side by side."""
    import ctypes as C
    L = host.lib()
    L.aqh_lane_schedule.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_double] + [C.c_void_p] * 4
    rng = np.random.default_rng(7)

    def schedule(R, W, cost, flags, gain=3.0):
        n = len(R)

        def csr(A):
            off = np.zeros(n + 1, np.int32)
            off[1:] = np.cumsum([len(a) for a in A])
            var = np.array([v for a in A for v, _ in a] + [0], np.int32)
            rows = np.array([m for a in A for _, m in a] + [0], np.uint32)
            return off, var, rows
        ro, rv, rr = csr(R)
        wo, wv, wr = csr(W)
        cost = np.asarray(cost, np.float64)
        flags = np.asarray(flags, np.uint8)
        lane, wait = np.zeros(n, np.int32), np.zeros(n, np.int32)
        marked = np.zeros(n, np.uint8)
        last1 = C.c_int(-2)
        p = lambda a: a.ctypes.data   # noqa: E731
        assert L.aqh_lane_schedule(n, p(ro), p(rv), p(rr), p(wo), p(wv), p(wr), p(cost), p(flags), gain, p(lane),
                                   p(wait), p(marked), C.addressof(last1)) == 0
        return lane, wait, marked, last1.value

    def conflict(a, b, R, W, flags):
        if (flags[a] | flags[b]) & 2:
            return True
        hit = lambda X, Y: any(v == w and (m & q) for v, m in X for w, q in Y)   # noqa: E731
        return hit(W[a], R[b]) or hit(W[a], W[b]) or hit(R[a], W[b])

    used_lane1 = 0
    for trial in range(300):
        n = int(rng.integers(2, 40))
        nv = int(rng.integers(2, 12))
        R, W, cost, flags = [], [], [], []
        for k in range(n):
            acc = lambda m: [(int(v), int(rng.choice([7, 7, 1, 2, 4, 5]))) for v in   # noqa: E731
                             rng.choice(nv, size=min(nv, int(rng.integers(0, m))), replace=False)]
            R.append(acc(4))
            W.append(acc(3))
            R[-1] += W[-1]                   # (an output may be read back)
            cost.append(float(rng.choice([1, 1, 1, 10, 10, 40])))
            f = 0
            if rng.random() < 0.15:
                f |= 1
            if rng.random() < 0.03:
                f |= 2
            if rng.random() < 0.15:
                f |= 4
            flags.append(f)
        lane, wait, marked, last1 = schedule(R, W, cost, flags, gain=float(rng.choice([0.0, 1.0, 3.0, 8.0])))
        live = [k for k in range(n) if not flags[k] & 4]
        assert all(lane[k] == 0 for k in range(n) if flags[k] & 3)
        on1 = [k for k in live if lane[k] == 1]
        assert last1 == (on1[-1] if on1 else -1) and (not on1 or marked[last1])
        used_lane1 += bool(on1)
        # happens-before: lane order + recorded events that somebody waits for
        hb = np.zeros((n, n), bool)
        for l in (0, 1):
            q = [k for k in live if lane[k] == l]
            for a, b in zip(q, q[1:]):
                hb[a, b] = True
        for k in live:
            if wait[k] >= 0:
                u = int(wait[k])
                assert u < k and lane[u] != lane[k] and marked[u] and not flags[u] & 4
                hb[u, k] = True
        for m in range(n):                   # transitive closure (indices are a topological order)
            hb |= np.outer(hb[:, m], hb[m, :])
        for b in live:
            for a in live:
                if a < b and conflict(a, b, R, W, flags):
                    assert hb[a, b], (trial, a, b, lane.tolist(), wait.tolist())
    assert used_lane1 > 100
    # two independent chains of three sweeps: side by side, no waits between them
    R = [[(0, 7)], [(0, 7)], [(0, 7)], [(1, 7)], [(1, 7)], [(1, 7)]]
    W = [[(0, 7)], [(0, 7)], [(0, 7)], [(1, 7)], [(1, 7)], [(1, 7)]]
    lane, wait, marked, last1 = schedule(R, W, [10] * 6, [0] * 6)
    assert lane.tolist() == [0, 0, 0, 1, 1, 1] and (wait == -1).all() and last1 == 5
