"""Generates tests/golden/kernel_signatures.json from the reference's OpenCL scripts: for every
__kernel of every .cl file under resources/Scripts and examples/**, the ordered argument list the
reference's Kernel tool would reflect with clGetKernelArgInfo (Kernel.cpp:497-556): name, whether
it is a pointer, and whether the pointer is const (`const` or `__constant`).  Run in the build
container only (/root/reference is not on the GPU box); the JSON is committed.

    python tests/golden/make_kernel_signatures.py"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_signatures.json")

MACROS = {
    "LINKLIST_LOCAL_PARAMS": "const __global usize * icell, const __global usize * ihoc, svec4 n_cells",
    "LINKLIST_REMOTE_PARAMS": "const __global usize * icell, const __global usize * mpi_icell, "
                              "const __global usize * mpi_ihoc, svec4 n_cells",
}


def strip_comments(t):
    t = re.sub(r"/\*.*?\*/", " ", t, flags=re.S)
    return re.sub(r"//[^\n]*", " ", t)


def parse_args(argtxt):
    for k, v in MACROS.items():
        argtxt = argtxt.replace(k, v)
    out = []
    for a in [x.strip() for x in argtxt.split(",") if x.strip()]:
        ptr = "*" in a
        const = bool(re.search(r"\bconst\b|__constant\b|\bconstant\b", a))
        name = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", a)[-1]
        out.append([name, ptr, const])
    return out


def main():
    sig = {}
    roots = [os.path.join(REF, "resources", "Scripts"), os.path.join(REF, "examples")]
    for root in roots:
        for d, _, files in os.walk(root):
            for f in sorted(files):
                if not f.endswith(".cl"):
                    continue
                path = os.path.join(d, f)
                txt = strip_comments(open(path, errors="replace").read())
                rel = os.path.relpath(path, os.path.join(REF, "resources", "Scripts")) if root.endswith("Scripts") \
                    else "examples:" + f
                # wrappers such as cfd/Shepard.cl define a macro and #include the script with the kernels
                for inc in re.findall(r'#include\s+"resources/Scripts/([^"]+\.cl)"', txt):
                    ip = os.path.join(REF, "resources", "Scripts", inc)
                    if os.path.exists(ip):
                        txt += "\n" + strip_comments(open(ip, errors="replace").read())
                for m in re.finditer(r"__kernel\s+void\s+([A-Za-z_0-9]+)\s*\((.*?)\)\s*\{", txt, flags=re.S):
                    sig.setdefault(rel, {})[m.group(1)] = parse_args(m.group(2))
    json.dump(sig, open(OUT, "w"), indent=0, sort_keys=True)
    print(len(sig), "scripts,", sum(len(v) for v in sig.values()), "kernels ->", OUT)


if __name__ == "__main__":
    main()
