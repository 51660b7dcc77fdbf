"""Builds libaquacuda.so (sm_100a only) in-tree with nvcc.

    python -m aquagpusph_b200.build [--force]

Object files go to build/, the library next to this file.  nvcc cross-compiles
without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libaquacuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
# per-file extra flags
EXTRA = {
    # bandwidth-bound kernels: keep the reference's uncontracted arithmetic
    "elementwise.cu": ["-fmad=false"],
}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "aquacuda.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [nvcc()] + ARCH + COMMON + EXTRA.get(src, []) + \
                ["-c", os.path.join(CSRC, src), "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
