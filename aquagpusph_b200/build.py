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
    headers.append(os.path.join(ROOT, "include", "aquasvm.h"))
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
    build_host(force, verbose)
    return LIB


HOST = os.path.join(HERE, "host")
HOSTLIB = os.path.join(HERE, "libaquahost.so")
HOSTEXE = os.path.join(HERE, "AQUAgpusph-b200")


def build_host(force=False, verbose=False):
    """C++ host (XML front-end, Variables, tools, scheduler) -> libaquahost.so + CLI."""
    os.makedirs(OBJ, exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    flags = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-I" + os.path.join(ROOT, "include"),
             "-I" + HOST]
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")]
    hdrs += [os.path.join(ROOT, "include", f) for f in ("aquacuda.h", "aquahost.h", "aquasvm.h")]
    hdrs.append(os.path.abspath(__file__))
    srcs = sorted(f for f in os.listdir(HOST) if f.endswith(".cpp") and f != "main.cpp")
    jobs, objs = [], []
    for src in srcs + ["main.cpp"]:
        obj = os.path.join(OBJ, "host_" + src[:-4] + ".o")
        if src != "main.cpp":
            objs.append(obj)
        if force or _stale(obj, [os.path.join(HOST, src)] + hdrs):
            jobs.append([cxx] + flags + ["-c", os.path.join(HOST, src), "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    link = ["-L" + HERE, "-laquacuda", "-Wl,-rpath,$ORIGIN"]
    if force or jobs or _stale(HOSTLIB, objs + [LIB]):
        run([cxx, "-shared", "-o", HOSTLIB] + objs + link)
    main_o = os.path.join(OBJ, "host_main.o")
    if force or jobs or _stale(HOSTEXE, [main_o, HOSTLIB]):
        run([cxx, "-o", HOSTEXE, main_o, "-L" + HERE, "-laquahost", "-laquacuda",
             "-Wl,-rpath,$ORIGIN"])
    return HOSTLIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
