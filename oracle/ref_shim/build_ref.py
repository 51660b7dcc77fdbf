#!/usr/bin/env python
"""TEST INFRASTRUCTURE: builds oracle/_ref/libaquaref{2,3}d.so from the reference's
OWN device scripts, read where they lie under /root/reference.

    python oracle/ref_shim/build_ref.py        (build container only)

For every hot-path script under resources/Scripts the text is passed through ONE
mechanical rewrite -- the OpenCL vector literal "(float4)(a, b)" becomes
"float4(a, b)" -- into a scratch directory under /tmp (never into this
repository), wrapped in a translation unit that includes cl_shim.hpp, and
compiled with g++ (-O2 -ffp-contract=off).  For every __kernel a driver
    extern "C" void aqref_<script>__<entry>(size_t n, <the kernel's own parameters>)
calls it once per work-item.  Outputs: only oracle/_ref/*.so (git-ignored).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AQUA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "_ref")

SCRIPTS = [
    "basic/time_scheme/euler.cl", "basic/time_scheme/improved_euler.cl",
    "basic/time_scheme/midpoint.cl", "basic/Domain.cl", "basic/Sort.cl", "basic/EOS.cl",
    "basic/Binormal.cl", "basic/neighs.cl", "basic/MLS.cl", "basic/SetBuffer.cl",
    "cfd/Interactions.cl", "cfd/Rates.cl", "cfd/TimeStep.cl", "cfd/Sensors.cl",
    "cfd/SensorsRenormalization.cl", "cfd/Shepard.cl", "cfd/deltaSPH.cl",
    "cfd/Boundary/BIe/Interactions.cl", "cfd/Boundary/BIe/Rates.cl",
    "cfd/Boundary/BIe/ElasticBounce.cl", "cfd/Boundary/BIe/PST.cl",
    "cfd/MPI.cl", "cfd/MPI/planes.cl",
    "cfd/Boundary/BI/LapU.cl", "cfd/Boundary/BI/GradP.cl", "cfd/Boundary/BI/Interpolation.cl",
    "cfd/Boundary/BI/InterpolationShepard.cl", "cfd/Boundary/BI/Interactions.cl",
    "cfd/Boundary/BI/Shepard.cl", "cfd/Boundary/ElasticBounce.cl",
    "cfd/Motions/Transform.cl", "cfd/Motions/UnTransform.cl", "cfd/Motions/Velocity.cl",
    "cfd/Motions/Acceleration.cl", "cfd/Energy/Energy.cl",
    "cfd/Energy/EnergyKin.cl", "cfd/Forces/Forces.cl", "basic/DensityClamp.cl", "basic/IdInverse.cl",
    "basic/time_scheme/adam_bashforth.cl", "cfd/Boundary/BI/NoSlip.cl",
    "cfd/Boundary/Symmetry/Mirror.cl",
    "cfd/ideal_gas/EOS.cl", "cfd/ideal_gas/Rates.cl", "cfd/ideal_gas/Sort.cl", "cfd/ideal_gas/TimeStep.cl",
    "cfd/ideal_gas/riemann/Rates.cl", "cfd/ideal_gas/time_scheme/midpoint.cl",
    "cfd/ideal_gas/symmetry/Mirror.cl", "cfd/ideal_gas/riemann/Interactions.cl",
    "cfd/ideal_gas/time_scheme/euler.cl", "cfd/ideal_gas/time_scheme/improved_euler.cl",
    "cfd/Boundary/Inlet/Inlet.cl", "cfd/Boundary/Outlet/Outlet.cl", "cfd/Boundary/Portal/Mirror.cl",
    "cfd/Boundary/Portal/Shepard.cl", "cfd/Boundary/Portal/Interactions.cl",
]
# basic/Shepard.cl and basic/deltaSPH.cl are compiled through their cfd/ wrappers

# Tool-layer kernels (embedded by the reference's build into its C++ tools, with the matching
# .hcl.in header prepended): only those without work-group cooperation can run one item at a
# time -- LinkList.cl.in (iHoc, iCell, linkList).  RadixSort.cl.in scans through __local memory
# between barriers and stays pinned by the reference's own property tests.
TOOL_SCRIPTS = [("aquagpusph/CalcServer/LinkList.cl.in", "aquagpusph/CalcServer/LinkList.hcl.in")]

# The same script under other compile-time definitions (the <Define>s of a case): (script, name it is
# indexed under, lines in front of the shim header)
VARIANTS = [("cfd/Interactions.cl", "cfd/Interactions@morris.cl", ["#define __LAP_FORMULATION__ 2"]),
            ("cfd/Boundary/Portal/Interactions.cl", "cfd/Boundary/Portal/Interactions@morris.cl",
             ["#define __LAP_FORMULATION__ 2"])]

VEC_LITERAL = re.compile(r"\(\s*(float2|float3|float4|float16|matrix|vec|vec2|vec3|vec4|vec_xyz)\s*\)\s*\(")
MACRO_PARAMS = {
    "LINKLIST_LOCAL_PARAMS": ["icell", "ihoc", "n_cells"],
    "LINKLIST_REMOTE_PARAMS": ["icell", "mpi_icell", "mpi_ihoc", "n_cells"],
}


def rewrite(text):
    return VEC_LITERAL.sub(lambda m: m.group(1) + "(", text)


def mirror(tmp):
    """Rewritten copy of resources/Scripts (headers and scripts) under tmp."""
    src = os.path.join(REF, "resources", "Scripts")
    for root, _, files in os.walk(src):
        for fn in files:
            if not fn.endswith((".cl", ".h", ".hcl")):
                continue
            p = os.path.join(root, fn)
            rel = os.path.relpath(p, REF)
            dst = os.path.join(tmp, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(p, encoding="utf-8", errors="replace") as f:
                txt = f.read()
            with open(dst, "w") as f:
                f.write(rewrite(txt))
    for files in TOOL_SCRIPTS:
        for rel in files:
            dst = os.path.join(tmp, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(os.path.join(REF, rel), encoding="utf-8", errors="replace") as f:
                txt = f.read()
            with open(dst, "w") as f:
                f.write(rewrite(txt))


def strip_comments(txt):
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return re.sub(r"//[^\n]*", "", txt)


def kernels_of(path, seen=None):
    """[(entry, [param text], [param names])] of a script, following its
    #include of another .cl (cfd/Shepard.cl -> basic/Shepard.cl)."""
    seen = seen or set()
    txt = strip_comments(open(path, encoding="utf-8", errors="replace").read())
    out = []
    for inc in re.findall(r'#include\s+"(resources/Scripts/[^"]+\.cl)"', txt):
        p = os.path.join(REF, inc)
        if p not in seen:
            seen.add(p)
            out += kernels_of(p, seen)
    for m in re.finditer(r"__kernel\s+void\s+(\w+)\s*\(([^)]*)\)", txt):
        params = [p.strip() for p in m.group(2).split(",") if p.strip()]
        names = []
        for p in params:
            if p in MACRO_PARAMS:
                names += MACRO_PARAMS[p]
            else:
                names.append(re.findall(r"(\w+)\s*$", p)[0])
        out.append((m.group(1), params, names))
    return out


def mangle(script):
    return re.sub(r"\W", "_", script[:-3])


def wrapper(script, dims, header=None, alias=None, predef=()):
    """header: a tool-layer script (path relative to the reference root) and the .hcl.in the
    reference's build prepends to it.  alias / predef: the same script under another name with
    other compile-time definitions (VARIANTS)."""
    tag = mangle(os.path.basename(script)[:-3] if header else (alias or script))
    ks = kernels_of(os.path.join(REF, script) if header else os.path.join(REF, "resources", "Scripts", script))
    lines = ["#define HAVE_%dD 1" % dims] + list(predef) + ['#include "cl_shim.hpp"']
    for entry, _, _ in ks:
        lines.append("#define %s aqrefk_%s__%s" % (entry, tag, entry))
    # the type headers define non-inline helpers (outer, det, inv): keep them TU-local
    lines.append("namespace {")
    if header:
        lines.append('#include "%s"' % header)
        lines.append('#include "%s"' % script)
    else:
        lines.append('#include "resources/Scripts/%s"' % script)
    lines.append("}")
    for entry, _, _ in ks:
        lines.append("#undef %s" % entry)
    for entry, params, names in ks:
        lines.append('extern "C" void aqref_%s__%s(size_t aqref_n, %s)\n{' % (tag, entry, ", ".join(params)))
        lines.append("    for (size_t aqref_g = 0; aqref_g < aqref_n; aqref_g++) { clshim_gid = aqref_g; aqrefk_%s__%s(%s); }\n}"
                     % (tag, entry, ", ".join(names)))
    out = []
    for entry, params, names in ks:
        kinds = []
        for p_ in params:
            if p_ in MACRO_PARAMS:
                kinds += ["ptr", "ptr", "ptr", "svec4"][-len(MACRO_PARAMS[p_]):] if p_ == "LINKLIST_REMOTE_PARAMS" \
                    else ["ptr", "ptr", "svec4"]
            elif "*" in p_:
                kinds.append("ptr")
            else:
                t = " ".join(p_.replace("const", "").split()[:-1])
                kinds.append({"float": "float", "usize": "uint", "uint": "uint", "unsigned int": "uint",
                              "int": "int", "vec": "vec", "svec4": "svec4", "uivec4": "svec4", "svec2": "svec2",
                              "vec4": "vec4"}[t.strip()])
        out.append((entry, names, kinds))
    return "\n".join(lines) + "\n", out


COMMON = r'''
#include "cl_shim.hpp"
thread_local size_t clshim_gid = 0;
clshim_defs clshim_D = { 1.f, 1.f, 1.f, 2.f, %d.f };
extern "C" void aqref_set_defs(float h_, float conw_, float conf_, float support_, float dims_)
{
    clshim_D.v_h = h_; clshim_D.v_conw = conw_; clshim_D.v_conf = conf_;
    clshim_D.v_support = support_; clshim_D.v_dims = dims_;
}
extern "C" int aqref_dims(void) { return %d; }
'''


def build(verbose=False):
    if not os.path.isdir(os.path.join(REF, "resources", "Scripts")):
        raise RuntimeError("reference tree not found at " + REF)
    os.makedirs(OUT, exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    flags = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w",
             "-I" + HERE]
    index = {}
    with tempfile.TemporaryDirectory(prefix="aqua_refshim_") as tmp:
        mirror(tmp)
        for dims in (2, 3):
            objs, jobs = [], []
            d = os.path.join(tmp, "tu%d" % dims)
            os.makedirs(d)
            src = os.path.join(d, "common.cpp")
            open(src, "w").write(COMMON % (dims, dims))
            jobs.append((src, src[:-4] + ".o"))
            for s in SCRIPTS:
                txt, ks = wrapper(s, dims)
                src = os.path.join(d, mangle(s) + ".cpp")
                open(src, "w").write(txt)
                jobs.append((src, src[:-4] + ".o"))
                index[s] = ks
            for s, alias, predef in VARIANTS:
                txt, ks = wrapper(s, dims, alias=alias, predef=predef)
                src = os.path.join(d, mangle(alias) + ".cpp")
                open(src, "w").write(txt)
                jobs.append((src, src[:-4] + ".o"))
                index[alias] = ks
            for s, hdr in TOOL_SCRIPTS:
                txt, ks = wrapper(s, dims, hdr)
                src = os.path.join(d, "tool_" + mangle(os.path.basename(s)[:-3]) + ".cpp")
                open(src, "w").write(txt)
                jobs.append((src, src[:-4] + ".o"))
                index[os.path.basename(s)[:-3]] = ks

            def cc(job):
                cmd = [cxx] + flags + ["-I" + tmp, "-c", job[0], "-o", job[1]]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode:
                    sys.stderr.write("FAILED: %s\n%s\n" % (" ".join(cmd), r.stderr[:4000]))
                    raise RuntimeError("ref_shim: cannot compile " + os.path.basename(job[0]))
                return job[1]

            with ThreadPoolExecutor(8) as ex:
                objs = list(ex.map(cc, jobs))
            lib = os.path.join(OUT, "libaquaref%dd.so" % dims)
            subprocess.check_call([cxx, "-shared", "-o", lib] + objs)
            if verbose:
                print(lib)
    # argument names of every kernel, for the Python caller (oracle/ref.py)
    with open(os.path.join(OUT, "kernels.txt"), "w") as f:
        for s, ks in sorted(index.items()):
            for entry, names, kinds in ks:
                f.write("%s %s %s %s\n" % (s, entry, ",".join(names), ",".join(kinds)))
    return OUT


if __name__ == "__main__":
    build(verbose=True)
