"""Which of the reference's shipped examples this build can run: every examples/{2D,3D}/*/src/templates
Main.xml is resolved by OUR front-end (tools/resolve_case.py) and each `kernel` tool's (script, entry)
is looked up in the CUDA registry; tool types the host does not provide are listed too.
Build container only (reads /root/reference).     python tools/example_coverage.py [--markdown]"""
import contextlib
import io
import os
import re
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import resolve_case as R  # noqa: E402
from aquagpusph_b200 import _lib  # noqa: E402

HOST_TYPES = {"kernel", "copy", "set", "set_scalar", "reduction", "link-list", "radix-sort", "sort", "unsort",
              "assert", "if", "while", "end", "endif", "mpi-sync", "dummy", "python", "installable", "report_screen",
              "report_file", "report_dump", "report_performance"}
DEFS = ("-DH=0.04f", "-DCONW=1.f", "-DCONF=1.f", "-DSUPPORT=2.f", "-DKERNEL_NAME=Wendland", "-D__LAP_MONAGHAN__=1",
        "-D__LAP_MORRIS__=2", "-D__LAP_FORMULATION__=__LAP_MONAGHAN__")


def runtime_scripts(missing, src, dims):
    """Of the (script, entry) pairs the registry lacks: those that the run-time script path compiles
    (aqc_script_check: NVRTC behind the dialect header, csrc/clc.cu), and those it does not."""
    root = tempfile.mkdtemp()
    os.makedirs(os.path.join(root, "resources"))
    os.symlink(os.path.join(R.REF, "resources", "Scripts"), os.path.join(root, "resources", "Scripts"))
    ok, bad = [], []
    for script, entry in missing:
        cands = [os.path.join(R.REF, "resources", "Scripts", script), os.path.join(R.REF, src, os.path.basename(script))]
        path = next((c for c in cands if os.path.exists(c)), None)
        try:
            if path is None:
                raise _lib.AquaError("no such file")
            _lib.script_check(path, entry, dims, root, ("-DDIMS=%d" % dims,) + DEFS)
            ok.append("%s::%s" % (script, entry))
        except _lib.AquaError:
            bad.append("%s::%s" % (script, entry))
    return ok, bad


def scan(jit=True):
    L = _lib.lib()
    R.OUT = tempfile.mkdtemp()
    rows = []
    for dims, D in ((2, "2D"), (3, "3D")):
        base = os.path.join(R.REF, "examples", D)
        for ex in sorted(os.listdir(base)):
            src = "examples/%s/%s/src/templates" % (D, ex)
            if not os.path.exists(os.path.join(R.REF, src, "Main.xml")):
                continue
            name = "%s_%s" % (ex, D)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    R.resolve(name, src, dims)
            except Exception:   # noqa: BLE001
                rows.append((D, ex, None, None, ["(Main.xml includes a file its generator writes)"], [], [],
                             ["(Main.xml includes a file its generator writes)"]))
                continue
            txt = open(os.path.join(R.OUT, name + ".xml")).read()
            ks = set(re.findall(r'type="kernel"[^>]*path="[^"]*?(?:Scripts/)?([^"]*\.cl)" entry_point="([^"]*)"', txt))
            miss = sorted(k for k in ks if L.aqc_kernel_lookup(("Scripts/" + k[0]).encode(), k[1].encode(), dims) < 0
                          and L.aqc_kernel_lookup(k[0].encode(), k[1].encode(), dims) < 0)
            other = sorted(set(re.findall(r'<Tool [^>]*type="([^"]*)"', txt)) - HOST_TYPES)
            # definitions the hand-written kernels do not honour (aqc_set_define refuses them at load): another
            # SPH kernel function than Wendland, another Laplacian than Monaghan's or Morris' (under the latter
            # cfd/Interactions.cl has a hand-written build, the other scripts with a Laplacian term run as scripts)
            for dname, ok in (("KERNEL_NAME", ("Wendland",)),
                              ("__LAP_FORMULATION__", ("__LAP_MONAGHAN__", "1", "__LAP_MORRIS__", "2"))):
                vals = re.findall(r'<Define name="%s" value="([^"]*)"' % dname, txt)
                if vals and vals[-1] not in ok:
                    other.append("definition %s=%s" % (dname, vals[-1]))
            jit_ok, jit_bad = runtime_scripts(miss, src, dims) if jit else ([], ["%s::%s" % k for k in miss])
            rows.append((D, ex, txt.count("<Tool "), len(ks), ["%s::%s" % k for k in miss], other, jit_ok, jit_bad))
    return rows


if __name__ == "__main__":
    rows = scan()
    if "--markdown" in sys.argv:
        print("| example | tools | kernels | scripts without a hand-written kernel (compiled at run time) | ... that do "
              "not compile | tool types / definitions not provided |")
        print("|---|---|---|---|---|---|")
        for D, ex, nt, nk, miss, other, jit_ok, jit_bad in rows:
            print("| %s/%s | %s | %s | %s | %s | %s |" % (D, ex, nt or "–", nk or "–",
                                                        ", ".join("`%s`" % m for m in jit_ok) or "none",
                                                        ", ".join("`%s`" % m for m in jit_bad) or "none",
                                                        ", ".join(other) or "none"))
    else:
        for r in rows:
            print(r)
    full = sum(1 for r in rows if r[2] and not r[4] and not r[5])
    runs = sum(1 for r in rows if r[2] and not r[7] and not r[5])
    print("\n%d of %d examples are covered by hand-written kernels alone, %d with run-time scripts" % (full, len(rows), runs))
