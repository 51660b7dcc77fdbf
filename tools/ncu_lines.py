"""Per-source-line instruction counts of a profiled kernel (ncu source page, needs -lineinfo).
    python tools/ncu_lines.py gpurun_out/prof_X.ncu-rep [top]"""
import csv, io, subprocess, sys, os
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fpath, per, tot = None, "?", [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name"):
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["Instructions Executed"]); t = int(d["Thread Instructions Executed"])
    except Exception:
        continue
    sm = d.get("# Samples", "0"); sm = int(sm) if sm.isdigit() else 0
    per.append((n, t, sm, fpath, r[0], r[1].strip()))
    tot += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total warp instructions", tot)
for n, t, sm, f, ln, src in sorted(per, key=lambda x: -x[0])[:top]:
    print("%-12s %4s inst %11d %5.1f%% thr/inst %4.1f smp %6d | %s" % (f, ln, n, 100.0 * n / tot, t / max(n, 1), sm, src[:90]))
