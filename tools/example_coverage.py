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
              "assert", "if", "while", "end", "endif", "mpi-sync", "dummy", "python", "report_screen",
              "report_file", "report_dump", "report_performance"}


def scan():
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
                rows.append((D, ex, None, None, ["(Main.xml includes a file its generator writes)"], []))
                continue
            txt = open(os.path.join(R.OUT, name + ".xml")).read()
            ks = set(re.findall(r'type="kernel"[^>]*path="[^"]*?(?:Scripts/)?([^"]*\.cl)" entry_point="([^"]*)"', txt))
            miss = sorted(k for k in ks if L.aqc_kernel_lookup(("Scripts/" + k[0]).encode(), k[1].encode(), dims) < 0
                          and L.aqc_kernel_lookup(k[0].encode(), k[1].encode(), dims) < 0)
            other = sorted(set(re.findall(r'<Tool [^>]*type="([^"]*)"', txt)) - HOST_TYPES)
            rows.append((D, ex, txt.count("<Tool "), len(ks), ["%s::%s" % k for k in miss], other))
    return rows


if __name__ == "__main__":
    rows = scan()
    if "--markdown" in sys.argv:
        print("| example | tools | kernels | scripts not in the CUDA registry | tool types not provided |")
        print("|---|---|---|---|---|")
        for D, ex, nt, nk, miss, other in rows:
            print("| %s/%s | %s | %s | %s | %s |" % (D, ex, nt or "–", nk or "–", ", ".join("`%s`" % m for m in miss) or "none",
                                                   ", ".join(other) or "none"))
    else:
        for r in rows:
            print(r)
    full = sum(1 for r in rows if r[2] and not r[4] and not r[5])
    print("\n%d of %d examples are covered completely" % (full, len(rows)))
