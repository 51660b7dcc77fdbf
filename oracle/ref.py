"""TEST INFRASTRUCTURE: the reference's OWN device scripts (resources/Scripts/**.cl)
compiled as C++ behind oracle/ref_shim/cl_shim.hpp into oracle/_ref/libaquaref{2,3}d.so
(built by oracle/ref_shim/build_ref.py in the build container, where /root/reference
exists; the .so files travel with the snapshot, the sources never enter the repo).

    R = ref.Ref(dims=3, h=0.03)
    R.run("cfd/Interactions.cl", "entry", n, V)     # V: dict name -> numpy array | scalar

Arguments are bound BY NAME exactly like the reference's Kernel tool does
(Kernel.cpp:497-556); one call per work-item, work-groups of one item with
LOCAL_MEM_SIZE defined (the '=' write-back variant the reference runs).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")


def available():
    return all(os.path.exists(os.path.join(_DIR, f))
               for f in ("libaquaref2d.so", "libaquaref3d.so", "kernels.txt"))


def build():
    """(Re)build from /root/reference; no-op when the tree is absent."""
    if os.path.isdir("/root/reference/resources/Scripts"):
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "build_ref", os.path.join(_HERE, "ref_shim", "build_ref.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.build()
    return available()


class _F2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class _F4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class _U4(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32), ("w", C.c_uint32)]


class _U2(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32)]


def _mangle(script):
    import re
    return re.sub(r"\W", "_", script[:-3])


class Ref:
    def __init__(self, dims, h):
        if not available():
            raise RuntimeError("oracle/_ref is not built (python oracle/ref_shim/build_ref.py)")
        self.dims = dims
        self.lib = C.CDLL(os.path.join(_DIR, "libaquaref%dd.so" % dims))
        assert self.lib.aqref_dims() == dims
        self.index = {}
        for line in open(os.path.join(_DIR, "kernels.txt")):
            s, e, names, kinds = line.split()
            self.index[(s, e)] = (names.split(","), kinds.split(","))
        self.set_h(h)

    def set_h(self, h):
        """basic.xml:119-123 evaluated like CalcServer.cpp:245-257 (6 significant digits)."""
        from . import oracle as O
        d = O.make_defs(self.dims, h)
        self.lib.aqref_set_defs(C.c_float(d.H), C.c_float(d.CONW), C.c_float(d.CONF),
                                C.c_float(d.SUPPORT), C.c_float(float(self.dims)))
        self.defs = d

    def kernels(self):
        return sorted(self.index)

    def run(self, script, entry, n, V, **over):
        names, kinds = self.index[(script, entry)]
        fn = getattr(self.lib, "aqref_%s__%s" % (_mangle(script), entry))
        fn.restype = None
        args = [C.c_size_t(int(n))]
        keep = []
        for name, kind in zip(names, kinds):
            v = over[name] if name in over else V[name]
            if kind == "ptr":
                a = v
                assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], name
                keep.append(a)
                args.append(C.c_void_p(a.ctypes.data))
            elif kind == "float":
                args.append(C.c_float(float(v)))
            elif kind == "uint":
                args.append(C.c_uint32(int(v)))
            elif kind == "int":
                args.append(C.c_int32(int(v)))
            elif kind == "vec":
                a = np.zeros(4, np.float32)
                vv = np.asarray(v, np.float32).ravel()
                a[:min(4, vv.size)] = vv[:4]
                args.append(_F4(*a) if self.dims == 3 else _F2(a[0], a[1]))
            elif kind == "vec4":
                a = np.asarray(v, np.float32).ravel()
                args.append(_F4(*a[:4]))
            elif kind == "svec4":
                a = np.asarray(v, np.uint32).ravel()
                args.append(_U4(*[int(x) for x in a[:4]]))
            elif kind == "svec2":
                a = np.asarray(v, np.uint32).ravel()
                args.append(_U2(int(a[0]), int(a[1])))
            else:
                raise ValueError(kind)
        fn(*args)
