"""A/B builds of the sweep engine: compiles csrc/sweeps.cu with extra -D flags and links it with
the other (already built) objects into build/variants/<name>/libaquacuda.so.  On the GPU box a
measurement script copies one over aquagpusph_b200/libaquacuda.so (the box's tree is scratch).

    python tools/build_variant.py <name> [-DS3_BATCH=1 ...]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aquagpusph_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
B.build()
out = os.path.join(B.ROOT, "build", "variants", name)
os.makedirs(out, exist_ok=True)
obj = os.path.join(out, "sweeps.o")
subprocess.check_call([B.nvcc()] + B.ARCH + B.COMMON + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "sweeps.cu"), "-o", obj],
                      stderr=open(os.path.join(out, "ptxas.log"), "w"))
objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.sources() if s != "sweeps.cu"] + [obj]
subprocess.check_call([B.nvcc()] + B.ARCH + ["-shared", "-o", os.path.join(out, "libaquacuda.so")] + objs)
print(os.path.join(out, "libaquacuda.so"))
