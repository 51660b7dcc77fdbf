"""SASS summary of the shipped library (profiles/r2_sass_summary.txt): cuobjdump -sass of
aquagpusph_b200/libaquacuda.so -> architectures, mnemonic totals, the Blackwell / Hopper-class instructions
and a per-kernel table.  No GPU needed.      python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aquagpusph_b200", "libaquacuda.so")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    names = {}
    try:
        dem = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True,
                             text=True, check=True).stdout.splitlines()
        for raw, d in zip(re.findall(r"Function : (\S+)", txt), dem):
            names[raw] = d
    except Exception:
        pass
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            total[op] += 1
            per[cur][op] += 1
            per[cur]["__n"] += 1
    n_ins = sum(c["__n"] for c in per.values())
    print("SASS of aquagpusph_b200/libaquacuda.so (cuobjdump -sass, tools/sass_summary.py): architectures %s, %d kernels, "
          "%d instructions\n" % (archs, len(per), n_ins))
    print("mnemonic totals (top 40):")
    for op, n in total.most_common(40):
        print("  %-12s %d" % (op, n))
    special = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MATCH", "REDUX", "ELECT", "UTMALDG", "UTCBAR", "HMMA",
               "UTCHMMA", "LDSM"]
    print("\nBlackwell / Hopper-class instructions (whole library): " +
          ", ".join("%s %d" % (k, total.get(k, 0)) for k in special))
    print("\nper kernel: instructions, FFMA, FFMA2, LDS, UBLKCP, SYNCS, MATCH, CALL (kernels with > 300 instructions, "
          "and the scalar-program kernel of the device loops)")
    for k, c in per.items():
        d = names.get(k, k)
        if c["__n"] > 300 or "svm_kernel" in d:
            print("  %6d %5d %5d %5d %4d %4d %4d %4d  %s" % (c["__n"], c["FFMA"], c["FFMA2"], c["LDS"], c["UBLKCP"],
                                                              c["SYNCS"], c["MATCH"], c["CALL"], d[:150]))


if __name__ == "__main__":
    sys.exit(main())
