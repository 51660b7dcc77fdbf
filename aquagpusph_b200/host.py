"""ctypes binding of libaquahost.so (include/aquahost.h): the C++ host that keeps
the reference's XML problem/tool API and runs it on libaquacuda.so.

    sim = Simulation("Main.xml", dims=3, root="/path/with/resources")
    sim.step(10); r = sim.download("r", unsorted=True); dt = sim.scalar("dt")
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libaquahost.so")
_lib = None

SYMBOLS = [
    "aqh_last_error", "aqh_set_log_level", "aqh_load", "aqh_parse", "aqh_destroy",
    "aqh_comm_unique_id", "aqh_comm_init",
    "aqh_write_resolved", "aqh_n_tools", "aqh_tool_name", "aqh_tool_type", "aqh_tool_elapsed_ms",
    "aqh_tool_used_times", "aqh_step", "aqh_run", "aqh_sync", "aqh_launch_count", "aqh_cuda_ctx",
    "aqh_fused_groups",
    "aqh_eval", "aqh_scalar_get", "aqh_scalar_set", "aqh_array_info", "aqh_array_download",
    "aqh_array_upload", "aqh_array_devptr",
]


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        from . import build as _build
        _build.build()
    # libaquacuda.so is found through the $ORIGIN rpath
    L = C.CDLL(_LIBPATH)
    L.aqh_last_error.restype = C.c_char_p
    L.aqh_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                           C.POINTER(C.c_void_p)]
    L.aqh_parse.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.aqh_destroy.argtypes = [C.c_void_p]
    L.aqh_comm_unique_id.argtypes = [C.c_void_p]
    L.aqh_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    L.aqh_write_resolved.argtypes = [C.c_void_p, C.c_char_p]
    L.aqh_n_tools.argtypes = [C.c_void_p]
    L.aqh_tool_name.argtypes = [C.c_void_p, C.c_int]
    L.aqh_tool_name.restype = C.c_char_p
    L.aqh_tool_type.argtypes = [C.c_void_p, C.c_int]
    L.aqh_tool_type.restype = C.c_char_p
    L.aqh_tool_elapsed_ms.argtypes = [C.c_void_p, C.c_int]
    L.aqh_tool_elapsed_ms.restype = C.c_double
    L.aqh_tool_used_times.argtypes = [C.c_void_p, C.c_int]
    L.aqh_tool_used_times.restype = C.c_uint
    L.aqh_step.argtypes = [C.c_void_p, C.c_int]
    L.aqh_run.argtypes = [C.c_void_p]
    L.aqh_sync.argtypes = [C.c_void_p]
    L.aqh_launch_count.argtypes = [C.c_void_p]
    L.aqh_launch_count.restype = C.c_uint64
    L.aqh_fused_groups.argtypes = [C.c_void_p]
    L.aqh_fused_groups.restype = C.c_uint
    L.aqh_cuda_ctx.argtypes = [C.c_void_p]
    L.aqh_cuda_ctx.restype = C.c_void_p
    L.aqh_eval.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.aqh_scalar_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.aqh_scalar_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.aqh_array_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]
    L.aqh_array_download.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    L.aqh_array_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    L.aqh_array_devptr.argtypes = [C.c_void_p, C.c_char_p]
    L.aqh_array_devptr.restype = C.c_void_p
    _lib = L
    return L


def _chk(rc):
    if rc:
        raise HostError(lib().aqh_last_error().decode())


def set_log_level(level):
    lib().aqh_set_log_level(int(level))


def evaluate(expr, type="float", decls="", dims=3, dtype=np.float32, n=1):
    """Variables::solve on the host: `decls` = "type name=value;..." (no device)."""
    out = np.zeros(n, dtype)
    _chk(lib().aqh_eval(dims, decls.encode(), type.encode(), expr.encode(), out.ctypes.data,
                        out.nbytes))
    return out[0] if n == 1 else out


def comm_unique_id():
    """ncclGetUniqueId on this process (rank 0); distribute the bytes to every rank."""
    buf = C.create_string_buffer(128)
    _chk(lib().aqh_comm_unique_id(buf))
    return buf.raw


class Simulation:
    """FileManager::load + CalcServer (parse_only=True: XML front-end only, no GPU)."""

    def __init__(self, xml_path, dims=3, device=-1, root=None, mpi_rank=0, mpi_size=1,
                 parse_only=False):
        self.h = C.c_void_p()
        self.dims = dims
        r = root.encode() if root else None
        if parse_only:
            _chk(lib().aqh_parse(xml_path.encode(), dims, r, C.byref(self.h)))
        else:
            _chk(lib().aqh_load(xml_path.encode(), dims, device, r, mpi_rank, mpi_size,
                                C.byref(self.h)))

    def comm_init(self, unique_id):
        """Join the run's NCCL communicator (unique_id: 128 bytes from comm_unique_id() of rank 0)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _chk(lib().aqh_comm_init(self.h, buf))

    def close(self):
        if self.h:
            lib().aqh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # pipeline
    def tools(self):
        L = lib()
        return [(L.aqh_tool_name(self.h, i).decode(), L.aqh_tool_type(self.h, i).decode())
                for i in range(L.aqh_n_tools(self.h))]

    def tool_times(self):
        L = lib()
        return [(L.aqh_tool_name(self.h, i).decode(), L.aqh_tool_used_times(self.h, i),
                 L.aqh_tool_elapsed_ms(self.h, i)) for i in range(L.aqh_n_tools(self.h))]

    def write_resolved(self, path):
        _chk(lib().aqh_write_resolved(self.h, path.encode()))

    def step(self, n=1):
        _chk(lib().aqh_step(self.h, int(n)))

    def run(self):
        _chk(lib().aqh_run(self.h))

    def sync(self):
        _chk(lib().aqh_sync(self.h))

    def launch_count(self):
        return int(lib().aqh_launch_count(self.h))

    def fused_groups(self):
        return int(lib().aqh_fused_groups(self.h))

    def cuda_ctx(self):
        return lib().aqh_cuda_ctx(self.h)

    # variables
    def scalar(self, name, dtype=np.float32, n=1):
        out = np.zeros(n, dtype)
        _chk(lib().aqh_scalar_get(self.h, name.encode(), out.ctypes.data, out.nbytes))
        return out[0] if n == 1 else out

    def set_scalar(self, name, expression):
        _chk(lib().aqh_scalar_set(self.h, name.encode(), str(expression).encode()))

    def array_info(self, name):
        n, eb = C.c_size_t(), C.c_size_t()
        _chk(lib().aqh_array_info(self.h, name.encode(), C.byref(n), C.byref(eb)))
        return int(n.value), int(eb.value)

    def has_array(self, name):
        n, eb = C.c_size_t(), C.c_size_t()
        return lib().aqh_array_info(self.h, name.encode(), C.byref(n), C.byref(eb)) == 0

    def download(self, name, dtype=np.float32, unsorted=False, out=None):
        n, eb = self.array_info(name)
        dt = np.dtype(dtype)
        ncomp = eb // dt.itemsize
        if out is None:
            out = np.empty((n, ncomp) if ncomp > 1 else (n,), dt)
        _chk(lib().aqh_array_download(self.h, name.encode(), out.ctypes.data, 1 if unsorted else 0))
        return out

    def upload(self, name, host):
        n, eb = self.array_info(name)
        host = np.ascontiguousarray(host)
        if host.nbytes != n * eb:
            raise HostError("upload(%s): %d bytes given, %d expected" % (name, host.nbytes, n * eb))
        _chk(lib().aqh_array_upload(self.h, name.encode(), host.ctypes.data))
