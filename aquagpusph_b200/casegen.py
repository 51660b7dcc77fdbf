"""Instantiate the resolved case templates (aquagpusph_b200/cases_xml/*.xml) for a
synthetic particle set and hand them to the C++ host.

The templates are the reference's example pipelines flattened by
tools/resolve_case.py; here the {{KEY}} placeholders of the example generators
(examples/*/src/Create.py `data = {...}`) are filled, the <Load>/<Save> file
references are dropped (arrays are uploaded from memory instead of FastASCII
files) and a Simulation is created.
"""
import os
import re
import tempfile

import numpy as np

from . import host

_HERE = os.path.dirname(os.path.abspath(__file__))
TEMPLATES = os.path.join(_HERE, "cases_xml")

STATE_FIELDS = ("r", "normal", "tangent", "u", "dudt", "rho", "drhodt", "m", "imove")


def _vec(v):
    return ", ".join(repr(float(x)) for x in v)


def instantiate(template, case, set_sizes, overrides=None, keep_reports=False):
    """Returns the XML text of `template` for the dict `case` (cases.py)."""
    txt = open(os.path.join(TEMPLATES, template + ".xml")).read()
    rep = {
        "DR": repr(case["dr"]), "HFAC": repr(case.get("hfac", case["h"] / case["dr"])),
        "CS": repr(case["cs"]), "COURANT": repr(case["courant"]),
        "DOMAIN_MIN": _vec(case["domain_min"]), "DOMAIN_MAX": _vec(case["domain_max"]),
        "REFD": repr(float(case["refd"][0])), "VISC_DYN": repr(float(case["visc_dyn"][0])),
        "DELTA": repr(float(case["delta"][0])), "G": repr(float(-case["g"][case["dims"] - 1]) + 0.0),
        "N": str(int(set_sizes[0])),
        "N_SENSORS": str(int(set_sizes[1]) if len(set_sizes) > 1 else 0),
        "NBC": str(int(set_sizes[1]) if len(set_sizes) > 1 else 0),
        "NFLUID": str(int(set_sizes[0])), "P0": repr(float(case.get("p0", 0.0))),
        "TEND": repr(float(case.get("t_end", 1.0))),
        "CLTYPE": "GPU", "CLDEVICE": "0", "CLPLATFORM": "0",
    }
    rep.update(case.get("placeholders", {}))   # the case's own keys (Create.py `data` of that example)
    for k, v in rep.items():
        txt = txt.replace("{{%s}}" % k, v)
    # the MPI example's Fluids.xml leaves the size of set 0 to the particle file
    txt = txt.replace("<ParticlesSet>", '<ParticlesSet n="%d">' % int(set_sizes[0]), 1)
    # particle data travel through aqh_array_upload, not through files
    txt = re.sub(r"\s*<Load [^>]*/>", "", txt)
    txt = re.sub(r"\s*<Save [^>]*/>", "", txt)
    if not keep_reports:
        # file/screen reports only cost host time; drop them like `-l 3` runs would hide them
        txt = re.sub(r"\s*<Report [^>]*/>", "", txt)
        txt = re.sub(r"\s*<Tool [^>]*type=\"report_(file|screen|performance)\"[^>]*/>", "", txt)
    for name, value in (overrides or {}).items():
        pat = r'(<Variable name="%s" [^>]*value=")[^"]*(")' % re.escape(name)
        if not re.search(pat, txt):
            raise KeyError("variable %s not found in template %s" % (name, template))
        txt = re.sub(pat, lambda m: m.group(1) + str(value) + m.group(2), txt)
    return txt


def load(template, case, set_sizes, overrides=None, device=0, workdir=None, keep_reports=False,
         mpi_rank=0, mpi_size=1, unique_id=None, transform=None, extra_tools=None):
    """Create a Simulation for `case` and upload its particle arrays."""
    txt = instantiate(template, case, set_sizes, overrides, keep_reports)
    if transform:
        txt = transform(txt)
    d = workdir or tempfile.mkdtemp(prefix="aqua_case_")
    path = os.path.join(d, "%s.rank%d.xml" % (template, mpi_rank))
    with open(path, "w") as f:
        f.write(txt)
    cwd = os.getcwd()
    os.chdir(d)  # report files are written relative to the working directory
    try:
        sim = host.Simulation(path, dims=case["dims"], device=device, mpi_rank=mpi_rank,
                              mpi_size=mpi_size)
    finally:
        os.chdir(cwd)
    for k in STATE_FIELDS:
        sim.upload(k, case[k])
    if mpi_size > 1:
        sim.comm_init(unique_id)
    sim.xml_path = path
    return sim


def move_tool(txt, name, after):
    """Re-place one <Tool> of a resolved XML right after the tool called `after`
    (what <Tool action="insert" after=...> would have produced)."""
    m = re.search(r'\n[ \t]*<Tool [^>]*name="%s" [^>]*?(/>|>.*?</Tool>)' % re.escape(name), txt, flags=re.S)
    if not m:
        raise KeyError("tool %s not found" % name)
    line = m.group(0)
    txt = txt[:m.start()] + txt[m.end():]
    a = re.search(r'\n[ \t]*<Tool [^>]*name="%s" [^>]*?(/>|>.*?</Tool>)' % re.escape(after), txt, flags=re.S)
    if not a:
        raise KeyError("tool %s not found" % after)
    return txt[:a.end()] + line + txt[a.end():]


def halo_mask_after_sort(txt):
    """The reference computes `mpi_neigh_mask` BEFORE the link-list sort of the same
    step (cfd/MPI/planes.xml:50 inserts it after "mpi remove") and uses it after
    the particles have been permuted, so the halo selection is stale whenever the
    sort moves particles (always on the first step).  Multi-device cases of this
    repository place the tool after the "Sort" stage instead; with that, N-device
    runs reproduce the single-device run to fp32 summation order."""
    names = re.findall(r'<Tool [^>]*name="([^"]*mpi neighs mask)"', txt)
    for nm in names:
        # keep the set_scalar tools that feed a prefixed plane (MPI.xml of the example)
        feeders = re.findall(r'<Tool [^>]*name="(%s mpi_plane_[rnp]\w*)"' % re.escape(nm.replace(" mask", "")), txt)
        anchor = "Sort"
        for f in feeders:
            txt = move_tool(txt, f, anchor)
            anchor = f
        txt = move_tool(txt, nm, anchor)
    return txt


def instantiate_plain(template, overrides=None, n=None):
    """A resolved template without placeholders (the reference's own test cases, e.g.
    tests/2D/MPI_plane): drops <Load>/<Save>/reports, optionally overrides variable
    values and the particle count."""
    txt = open(os.path.join(TEMPLATES, template + ".xml")).read()
    txt = re.sub(r"\s*<Load [^>]*/>", "", txt)
    txt = re.sub(r"\s*<Save [^>]*/>", "", txt)
    txt = re.sub(r"\s*<Report [^>]*/>", "", txt)
    txt = re.sub(r"\s*<Tool [^>]*type=\"report_(file|screen|performance)\"[^>]*/>", "", txt)
    if n is not None:
        txt = re.sub(r'(<ParticlesSet n=")\d+(")', lambda m: m.group(1) + str(int(n)) + m.group(2), txt)
    for name, value in (overrides or {}).items():
        pat = r'(<Variable name="%s" [^>]*value=")[^"]*(")' % re.escape(name)
        hits = list(re.finditer(pat, txt))
        if not hits:
            raise KeyError("variable %s not found in template %s" % (name, template))
        m = hits[-1]    # the last definition wins (Variable.cpp:1058-1064)
        txt = txt[:m.start()] + m.group(1) + str(value) + m.group(2) + txt[m.end():]
    return txt


def load_plain(template, arrays, dims, overrides=None, device=0, mpi_rank=0, mpi_size=1, n=None,
               unique_id=None):
    """Simulation for a placeholder-free template; `arrays` = {variable: host array}."""
    txt = instantiate_plain(template, overrides, n)
    d = tempfile.mkdtemp(prefix="aqua_case_")
    path = os.path.join(d, "%s.rank%d.xml" % (template, mpi_rank))
    with open(path, "w") as f:
        f.write(txt)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        sim = host.Simulation(path, dims=dims, device=device, mpi_rank=mpi_rank, mpi_size=mpi_size)
    finally:
        os.chdir(cwd)
    for k, a in arrays.items():
        sim.upload(k, a)
    if mpi_size > 1:
        sim.comm_init(unique_id)
    sim.xml_path = path
    return sim


def add_tool_after(txt, after, tool_xml):
    """Insert one <Tool .../> line after the tool called `after` in a resolved XML."""
    a = re.search(r'\n[ \t]*<Tool [^>]*name="%s" [^>]*?(/>|>.*?</Tool>)' % re.escape(after), txt, flags=re.S)
    if not a:
        raise KeyError("tool %s not found" % after)
    return txt[:a.end()] + "\n        " + tool_xml + txt[a.end():]


def multi_device_fixes(txt):
    """What this repository changes in the reference's MPI example pipeline so that an
    N-device run is a consistent simulation: the halo mask is taken after
    the sort (halo_mask_after_sort), dt / the midpoint residual are all-reduced, so
    every rank advances with the same time step and leaves the inner loop together, and
    `remove` sees the outgoing mask ((iii) below).
    When the halo exchange sits outside the midpoint loop (the example's include order,
    TODO at cfd/MPI.xml:21-23) it is moved inside, right after "midpoint eos": the
    reference exchanges u and rho of the halo once per step and iterates on stale
    copies, which makes the N-device result differ from the 1-device one by ~1e-2 next
    to the cuts; refreshed every sub-iteration the two agree to fp32 summation order."""
    txt = halo_mask_after_sort(txt)
    order = [m.group(1) for m in re.finditer(r'<Tool [^>]*name="([^"]*)"', txt)]
    # (iii) cfd/MPI.cl::remove (:176-199) decides by mpi_local_mask which LOCAL rows have left, but
    # mpi-sync has rewritten that mask by then (MPISync.cpp:222-223: own rank everywhere, the
    # sender's rank over what arrived): rows 0 .. n_received-1 are parked instead of the rows that
    # were sent, so every migration duplicates a particle on the receiver and destroys an
    # unrelated one on the sender.  The outgoing mask is kept and handed back for `remove`.
    if "mpi local sync" in order and "mpi remove" in order and "mpi append" in order:
        txt = txt.replace("    </Variables>",
                          '        <Variable name="mpi_sent_mask" type="size_t*" length="n_radix" />\n'
                          "    </Variables>", 1)
        txt = add_tool_after(txt, "mpi copy", '<Tool action="add" name="mpi sent mask backup" type="copy" '
                             'once="false" in="mpi_local_mask" out="mpi_sent_mask" />')
        txt = add_tool_after(txt, "mpi append", '<Tool action="add" name="mpi sent mask restore" type="copy" '
                             'once="false" in="mpi_sent_mask" out="mpi_local_mask" />')
    if "midpoint eos" in order and "mpi neighs sync" in order and \
            order.index("mpi neighs sync") < order.index("midpoint loop"):
        chain = ["mpi neighs mask reset"]
        for nm in [n for n in order if n.endswith("mpi neighs mask")]:
            chain += [n for n in order if re.match(re.escape(nm.replace(" mask", "")) + r" mpi_plane_", n)]
            chain.append(nm)
        chain += ["mpi neighs copy", "mpi neighs sync"]
        anchor = "midpoint eos"
        for nm in chain:
            txt = move_tool(txt, nm, anchor)
            anchor = nm
        # r and imove are fixed inside the loop (midpoint.cl:93-111 advances u and rho), and the
        # plane masks are functions of them alone (cfd/MPI/planes.cl:41-99): every sub-iteration
        # after the first reuses the first one's sort and counts (MPISync `depends`, ours)
        txt = re.sub(r'(<Tool [^>]*name="mpi neighs sync" [^>]*?)(\s*/>)',
                     lambda m: m.group(1) + ' depends="r,imove"' + m.group(2), txt, 1)
        # ... and so does the halo link-list: the positions it hashes are the neighbours' r, fixed
        # on every rank at once (link-list `depends`, ours)
        txt = re.sub(r'(<Tool [^>]*name="mpi link-list" [^>]*?)(\s*/>)',
                     lambda m: m.group(1) + ' depends="r,imove"' + m.group(2), txt, 1)
    txt = add_tool_after(txt, "cfd minimum time step",
                         '<Tool action="add" name="mpi global dt" type="mpi-allreduce" once="false" '
                         'in="dt" operation="min" />')
    if "midpoint residual" in order:
        txt = add_tool_after(txt, "midpoint residual",
                             '<Tool action="add" name="mpi global residual" type="mpi-allreduce" '
                             'once="false" in="Residual_midpoint" operation="sum" />')
    return txt


def _tool_text(txt, name):
    m = re.search(r'\n[ \t]*(<Tool [^>]*name="%s" [^>]*?(/>|>.*?</Tool>))' % re.escape(name), txt, flags=re.S)
    if not m:
        raise KeyError("tool %s not found" % name)
    return m.group(1)


def slab_delta_sph(txt, delta=0.1):
    """The slab pipeline of the reference's MPI example (after multi_device_fixes) with the
    delta-SPH and MLS stages of the single-device dam break (examples/3D/spheric_testcase2_dambreak:
    presets cfd/deltaSPH-full.xml + basic/MLS.xml): what BASELINE config 2 runs on one GPU, on N.

    The reference cannot do this: its MPI preset exchanges r, u, rho, m of the halo and adds the
    remote terms of cfd/Interactions.cl and the Shepard factor only (cfd/MPI.xml:59-87).  Here
      * the halo is exchanged once more BEFORE the midpoint loop (the MLS matrix is computed there,
        basic/MLS.xml), which also builds the halo link-list for the whole step: positions are
        fixed inside the loop, so the in-loop copy of that tool goes;
      * aqua/MPIdeltaSPH.cl::mls / ::full_lapp / ::lapp_corr (ours) add the remote terms of
        basic/MLS.cl, deltaSPH.cl::full + ::lapp and ::lapp_corr right after their local twins;
      * the corrected gradient lap_p_corr of the halo particles, which lapp_corr needs, travels in
        a second mpi-sync per sub-iteration on the OUTGOING mask of the first (kept in
        mpi_sent_neigh_mask), and is put in the order of the halo list like the other fields."""
    src = open(os.path.join(TEMPLATES, "spheric2_dambreak_3d.xml")).read()
    variables = (
        '        <Variable name="delta" type="float*" length="n_sets" />\n'
        '        <Variable name="lap_p" type="float*" length="N" />\n'
        '        <Variable name="mls_imove" type="unsigned int" value="1" />\n'
        '        <Variable name="mls" type="matrix*" length="N" />\n'
        '        <Variable name="mls_fluid" type="matrix*" length="N" />\n'
        '        <Variable name="lap_p_corr" type="vec*" length="N" />\n'
        '        <Variable name="mpi_lap_p_corr" type="vec*" length="n_radix" />\n'
        '        <Variable name="mpi_lap_p_corr_in" type="vec*" length="n_radix" />\n'
        '        <Variable name="mpi_neigh_mask2" type="size_t*" length="n_radix" />\n'
        '        <Variable name="mpi_sent_neigh_mask" type="size_t*" length="n_radix" />\n')
    txt = txt.replace("    </Variables>", variables + "    </Variables>", 1)
    # the definitions the single-device example adds to the presets' (the delta-SPH switches and
    # __DR_FACTOR__ = 0.75f, the element radius of BIe/ElasticBounce.cl and BIe/PST.cl: without it
    # particles next to a wall bounce at another distance than on one device)
    have = set(re.findall(r'<Define name="([^"]*)"', txt))
    last = list(re.finditer(r'\n[ \t]*<Define [^>]*/>', txt))[-1]
    extra = "".join(m.group(0) for m in re.finditer(r'\n[ \t]*<Define name="([^"]*)"[^>]*/>', src)
                    if m.group(1) not in have)
    txt = txt[:last.end()] + extra + txt[last.end():]
    txt = re.sub(r'(<Scalar name="visc_dyn" [^>]*/>)', lambda m: m.group(1) +
                 '\n        <Scalar name="delta" value="%r" />' % float(delta), txt)

    def tool(name, typ, attrs):
        return '<Tool action="add" name="%s" type="%s" once="false" %s />' % (name, typ, attrs)

    def kernel(name, entry, n=""):
        return tool(name, "kernel", 'path="Scripts/aqua/MPIdeltaSPH.cl" entry_point="%s" n="%s"' % (entry, n))

    order = [m.group(1) for m in re.finditer(r'<Tool [^>]*name="([^"]*)"', txt)]
    # ---- before the loop: halo exchange, halo link-list, MLS with its remote term
    chain = [n for n in order if order.index("midpoint eos") < order.index(n) <= order.index("mpi neighs sync")]
    pre = [tool("cfd reinit lap_p_corr", "set", 'in="lap_p_corr" value="VEC_ZERO"'), _tool_text(src, "MLS_imove1")]
    for nm in chain:
        pre.append(_tool_text(txt, nm).replace('name="%s"' % nm, 'name="pre %s"' % nm).replace(' depends="r,imove"', ""))
    for nm in ("mpi backup r", "mpi backup iset", "mpi backup u", "mpi backup rho", "mpi backup m"):
        pre.append(_tool_text(txt, nm).replace('name="%s"' % nm, 'name="pre %s"' % nm))
    pre.append(_tool_text(txt, "mpi link-list").replace(' depends="r,imove"', ""))
    pre.append(_tool_text(txt, "mpi sort").replace('name="mpi sort"', 'name="pre mpi sort"'))
    pre += [_tool_text(src, "imove1_MLS interactions"), kernel("mpi mls", "mls"), _tool_text(src, "imove1_MLS"),
            _tool_text(src, "MLS_mls_fluid"), _tool_text(src, "cfd reinit lap_p")]
    txt = re.sub(r'\n[ \t]*<Tool [^>]*name="mpi link-list" [^>]*?/>', "", txt, 1)   # (once per step is enough)
    anchor = "Sort"
    for t_xml in pre:
        txt = add_tool_after(txt, anchor, t_xml)
        anchor = re.search(r'name="([^"]*)"', t_xml).group(1)
    # ---- inside the loop
    txt = add_tool_after(txt, "mpi neighs copy", tool("mpi sent neigh mask backup", "copy",
                                                     'in="mpi_neigh_mask" out="mpi_sent_neigh_mask"'))
    anchor = "cfd interactions"
    for nm in ("cfd lap p mls", "cfd lap p full", "cfd lap p"):
        txt = add_tool_after(txt, anchor, _tool_text(src, nm))
        anchor = nm
    txt = add_tool_after(txt, "mpi shepard", kernel("mpi lap p", "full_lapp"))
    anchor = "Interactions"
    for t_xml in (_tool_text(src, "cfd lap p correction mls"),
                  kernel("mpi g copy", "copy_g"),
                  tool("mpi g mask", "copy", 'in="mpi_sent_neigh_mask" out="mpi_neigh_mask2"'),
                  tool("mpi g sync", "mpi-sync", 'mask="mpi_neigh_mask2" fields="mpi_lap_p_corr" processes="" '
                                                 'depends="r,imove"'),
                  tool("mpi g backup", "copy", 'in="mpi_lap_p_corr" out="mpi_lap_p_corr_in"'),
                  kernel("mpi g sort", "sort_g"),
                  _tool_text(src, "cfd lap p apply correction"),
                  kernel("mpi lap p apply correction", "lapp_corr"),
                  _tool_text(src, "LapP Correction")):
        txt = add_tool_after(txt, anchor, t_xml)
        anchor = re.search(r'name="([^"]*)"', t_xml).group(1)
    txt = add_tool_after(txt, "cfd rates", _tool_text(src, "cfd delta-SPH"))
    return txt


def slab_fixes_delta_sph(delta):
    """Transform for casegen.load: multi_device_fixes + slab_delta_sph."""
    return lambda txt: slab_delta_sph(multi_device_fixes(txt), delta)


def spheric2_slab(n_total, rank, size, hfac=3.0, overrides=None, device=0, unique_id=None, seed=None,
                  jitter=0.0, uscale=0.1, delta_sph=False, **kw):
    """BASELINE config 3: the 3-D dam break on `size` devices (y slabs) through the
    pipeline of examples/3D/spheric_testcase2_dambreak_mpi (131 tools: midpoint, BIe
    boundaries, variable time step, cfd/MPI.xml migration + halo)."""
    from . import cases
    c = cases.spheric2_dam_break_slab(n_total, hfac, rank, size, seed=seed, jitter=jitter, uscale=uscale)
    sim = load("spheric2_dambreak_mpi_3d", c, (c["n_set0"], c["N"] - c["n_set0"]), overrides, device,
               mpi_rank=rank, mpi_size=size, unique_id=unique_id,
               transform=slab_fixes_delta_sph(float(c["delta"][0])) if delta_sph else multi_device_fixes, **kw)
    return sim, c


def prescribed_roll(theta0=0.0698, period=1.94):
    """Transform for the tuned-liquid-damper template: the two `python` tools of cfd/motion.xml
    become `set_scalar` tools.

    `cfd motion data` (examples/2D/spheric_testcase9_tld/src/templates/Motion.py:154-185) integrates a
    mechanical model driven by the recorded mass position `T_1-94_A100mm_water.dat`, which the
    reference tree does not ship: it is replaced by the prescribed roll
    theta(t) = theta0 sin(2 pi t / period) with its two derivatives, evaluated at the same `t` the
    script reads.  `cfd motion state` (resources/Scripts/cfd/Motions/State.py:30-57) is reproduced
    exactly, including its first call, which hands the NEW state to UnTransform as the one to undo."""
    w = "(2 * pi / motion_period)"
    ph = "(2 * pi * t / motion_period)"
    variables = (
        '        <Variable name="motion_theta0" type="float" value="%r" />\n'
        '        <Variable name="motion_period" type="float" value="%r" />\n'
        '        <Variable name="motion_first" type="unsigned int" value="1" />\n'
        '        <Variable name="motion_r_bak" type="vec" value="0.0, 0.0, 0.0, 0.0" />\n'
        '        <Variable name="motion_a_bak" type="vec4" value="0.0, 0.0, 0.0, 0.0" />\n' % (theta0, period))

    def tool(name, var, value):
        return ('        <Tool action="add" name="%s" type="set_scalar" once="false" in="%s" value="%s" />\n'
                % (name, var, value.replace("<", "&lt;")))

    def keep(var, comps):
        return ", ".join("motion_first ? %s_%s : %s_bak_%s" % (var, c, var, c) for c in comps)

    data = (tool("cfd motion data", "motion_a", "0, 0, motion_theta0 * sin%s, 0" % ph) +
            tool("cfd motion data dadt", "motion_dadt", "0, 0, motion_theta0 * %s * cos%s, 0" % (w, ph)) +
            tool("cfd motion data ddaddt", "motion_ddaddt",
                 "0, 0, 0 - motion_theta0 * %s * %s * sin%s, 0" % (w, w, ph)))

    def state(dims):
        rc = "xy" if dims == 2 else "xyzw"
        return (tool("cfd motion state", "motion_r_bak", keep("motion_r", rc)) +
                tool("cfd motion state a", "motion_a_bak", keep("motion_a", "xyzw")) +
                tool("cfd motion state started", "motion_first", "0") +
                tool("cfd motion state r_in", "motion_r_in", ", ".join("motion_r_bak_" + c for c in rc)) +
                tool("cfd motion state a_in", "motion_a_in", ", ".join("motion_a_bak_" + c for c in "xyzw")) +
                tool("cfd motion state backup r", "motion_r_bak", ", ".join("motion_r_" + c for c in rc)) +
                tool("cfd motion state backup a", "motion_a_bak", ", ".join("motion_a_" + c for c in "xyzw")))

    def transform(txt, dims=2):
        for name, new in (("cfd motion data", data), ("cfd motion state", state(dims))):
            pat = r'[ \t]*<Tool [^>]*name="%s" type="python"[^>]*/>\n' % name
            if not re.search(pat, txt):
                raise KeyError("python tool '%s' not found" % name)
            txt = re.sub(pat, lambda m: new, txt, 1)
        return txt.replace("    </Variables>", variables + "    </Variables>", 1)

    return transform


SCRIPTS = os.path.join(TEMPLATES, "scripts")


def python_roll(theta0=0.0698, period=1.94):
    """Transform for the tuned-liquid-damper template that KEEPS the two `python` tools of
    cfd/motion.xml and points them at this repository's scripts (cases_xml/scripts/PrescribedRoll.py,
    MotionState.py) -- the same prescribed roll as `prescribed_roll`, through the python-tool route
    (host: aqh_set_script_runner + aquagpusph_b200/pytool.py)."""
    variables = ('        <Variable name="motion_theta0" type="float" value="%r" />\n'
                 '        <Variable name="motion_period" type="float" value="%r" />\n' % (theta0, period))

    def transform(txt, dims=2):
        for name, script in (("cfd motion data", "PrescribedRoll.py"), ("cfd motion state", "MotionState.py")):
            pat = r'(<Tool [^>]*name="%s" type="python"[^>]*path=")[^"]*(")' % name
            if not re.search(pat, txt):
                raise KeyError("python tool '%s' not found" % name)
            txt = re.sub(pat, lambda m: m.group(1) + os.path.join(SCRIPTS, script) + m.group(2), txt, 1)
        return txt.replace("    </Variables>", variables + "    </Variables>", 1)

    return transform


def spheric9_tld(n=10000, hfac=4.0, overrides=None, device=0, seed=None, theta0=0.0698, period=1.94, **kw):
    """BASELINE config 4 (2-D SPHERIC test 9, tuned liquid damper) through the 104-tool pipeline of
    examples/2D/spheric_testcase9_tld with the prescribed roll of `prescribed_roll`."""
    from . import cases
    c = cases.spheric9_tld_2d(n, hfac, seed=seed)
    sim = load("spheric9_tld_2d", c, (c["n_set0"], c["n_set1"]), overrides, device,
               transform=prescribed_roll(theta0, period), **kw)
    return sim, c


def lattice(n_side=100, hfac=2.0, overrides=None, device=0, **kw):
    """BASELINE config 5 (uniform 3-D lattice, all fluid, g = 0) through the 36-tool pipeline of
    cases_xml/src/lattice_3d/Lattice.xml: the reference's presets basic + improved Euler + cfd +
    variableTimeStep, i.e. predictor, link-list, sort, EOS, Shepard + Interactions, Rates, corrector,
    per-particle time step + min reduction."""
    from . import cases
    c = cases.lattice(n_side, hfac)
    c["courant"] = 0.25
    sim = load("lattice_3d", c, (c["N"],), overrides, device, **kw)
    return sim, c


def lattice_slab(n_side, rank, size, hfac=2.0, overrides=None, device=0, unique_id=None, nz_local=0, **kw):
    """BASELINE config 5 on `size` devices: z slabs of the lattice through the 76-tool pipeline of
    cases_xml/src/lattice_mpi_3d (the lattice pipeline + the reference's cfd/MPI.xml migration and
    halo presets), with the multi-device additions of `multi_device_fixes`.  nz_local = 0: the
    n_side^3 lattice cut in `size` slabs (strong scaling, the parity tests); nz_local > 0: every rank
    generates its own n_side x n_side x nz_local block (weak scaling, the bench)."""
    from . import cases
    if nz_local:
        c = cases.lattice_slab_local(n_side, nz_local, hfac, rank, size)
    else:
        c = cases.lattice_slab(n_side, hfac, rank, size)
    sim = load("lattice_mpi_3d", c, (c["N"],), overrides, device, mpi_rank=rank, mpi_size=size,
               unique_id=unique_id, transform=multi_device_fixes, **kw)
    return sim, c


def spheric3_lid_driven(nx=200, hfac=4.0, overrides=None, device=0, **kw):
    """The lid-driven cavity (SPHERIC test 3) through the unchanged 55-tool pipeline of
    examples/2D/spheric_testcase3_liddriven (improved Euler, delta-SPH full, BI boundaries, BINoSlip)."""
    from . import cases
    c = cases.spheric3_lid_driven_2d(nx, hfac)
    sim = load("spheric3_liddriven_2d", c, (c["n_set0"], c["n_set1"]), overrides, device, **kw)
    return sim, c


def souto2012_standing_wave(ny=100, hfac=4.0, overrides=None, device=0, **kw):
    """The standing wave of examples/2D/souto_etal_2012_standingwave through its unchanged 91-tool pipeline
    (improved Euler, delta-SPH full, BI bottom, elastic bounce, TWO symmetry planes of cfd/symmetry.xml feeding
    on buffer particles, kinetic-energy report)."""
    from . import cases
    c = cases.souto2012_standing_wave_2d(ny, hfac)
    sim = load("souto2012_standingwave_2d", c, (c["N"],), overrides, device, **kw)
    return sim, c


def shock_point(n=50000, hfac=2.0, rim_script="bc.cl", overrides=None, device=0, **kw):
    """The circular blast of examples/2D/shock_point through its unchanged 73-tool pipeline (midpoint scheme
    with autostop / autorelax, the cfd presets, the ideal-gas EOS / energy rates / energy time scheme, and a
    case-local script that freezes the rim: `rim_script` = the file the two `bc.cl` tools should read --
    compiled at run time; needs AQUAGPUSPH_ROOT to hold resources/Scripts/types/types.h for its include)."""
    from . import cases
    c = cases.shock_point_2d(n, hfac)
    tr = kw.pop("transform", None)

    def transform(txt):
        txt = txt.replace('path="bc.cl"', 'path="%s"' % rim_script)
        return tr(txt) if tr else txt
    sim = load("shock_point_2d", c, (c["N"],), overrides, device, transform=transform, **kw)
    for k in ("eint", "deintdt"):
        sim.upload(k, c[k])
    return sim, c


def spheric2(n=100000, hfac=3.0, overrides=None, device=0, seed=None, jitter=0.0, uscale=0.1, **kw):
    """BASELINE config 2 (3-D SPHERIC test 2 dam break) through the unchanged
    116-tool pipeline of examples/3D/spheric_testcase2_dambreak."""
    from . import cases
    c = cases.spheric2_dam_break(n, hfac, seed=seed, jitter=jitter, uscale=uscale)
    sim = load("spheric2_dambreak_3d", c, (c["N"] - 8, 8), overrides, device, **kw)
    return sim, c
