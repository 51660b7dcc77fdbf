"""ctypes binding of libaquacuda.so (the C-ABI in include/aquacuda.h).

This is the thin Python host used by the tests and bench.py; the C++ host
(aquagpusph_b200/host) binds the same symbols.  There is no CPU fallback: if the
library is missing it is built with nvcc, and creating a Context without a CUDA
device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libaquacuda.so")
_lib = None

# every symbol include/aquacuda.h declares (checked by tests/test_host_cpu.py)
SYMBOLS = [
    "aqc_ctx_create", "aqc_ctx_destroy", "aqc_last_error", "aqc_set_stream", "aqc_get_stream",
    "aqc_sync", "aqc_launch_count", "aqc_device_sm_count", "aqc_define_round6", "aqc_set_defs",
    "aqc_set_define",
    "aqc_alloc", "aqc_free", "aqc_host_alloc", "aqc_host_free", "aqc_memcpy_h2d",
    "aqc_memcpy_d2h", "aqc_memcpy_d2d", "aqc_side_fork", "aqc_memcpy_d2h_side", "aqc_side_record", "aqc_side_wait", "aqc_fill", "aqc_linklist_build", "aqc_radix_sort",
    "aqc_scatter_fields", "aqc_reduce", "aqc_kernel_lookup", "aqc_kernel_count",
    "aqc_kernel_name", "aqc_kernel_nargs", "aqc_kernel_args", "aqc_launch", "aqc_script_compile", "aqc_script_check", "aqc_event_create",
    "aqc_event_destroy", "aqc_event_record", "aqc_event_sync", "aqc_event_elapsed_ms",
    "aqc_comm_unique_id", "aqc_comm_init", "aqc_comm_destroy", "aqc_comm_rank", "aqc_comm_size",
    "aqc_mpi_sync", "aqc_mpi_sync_plan", "aqc_mpi_sync_ex", "aqc_mpi_sync_stats", "aqc_allreduce", "aqc_allreduce_host", "aqc_fused_lookup", "aqc_launch_fused",
    "aqc_fused_prefix", "aqc_kernel_write_rows", "aqc_kernel_read_rows", "aqc_fused_read_rows", "aqc_sweep_engine_select",
    "aqc_pairs_cache_enable", "aqc_pairs_cache_invalidate", "aqc_pairs_cache_stats", "aqc_pairs_cache_stats_remote", "aqc_fp32_peak",
    "aqc_watch_create", "aqc_watch_dirty", "aqc_watch_reset",
    "aqc_loop_create", "aqc_loop_destroy", "aqc_loop_table", "aqc_loop_begin", "aqc_loop_svm",
    "aqc_loop_end", "aqc_loop_abort", "aqc_loop_run", "aqc_loop_start", "aqc_loop_stats", "aqc_kernel_dev_scalars",
    "aqc_launch_ex", "aqc_lane_select", "aqc_lane_event", "aqc_lane_wait",
]

OP_SUM, OP_MIN, OP_MAX = 0, 1, 2
T_F32, T_U32, T_I32, T_VEC2, T_VEC4 = 0, 1, 2, 3, 4
ARG_ARRAY_IN, ARG_ARRAY_OUT, ARG_SCALAR, ARG_ARRAY_RO = 0, 1, 2, 3


class Defs(C.Structure):
    _fields_ = [("dims", C.c_int), ("H", C.c_float), ("CONW", C.c_float), ("CONF", C.c_float),
                ("SUPPORT", C.c_float), ("DIMS", C.c_float)]


class ArgInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_char_p), ("kind", C.c_int)]


class AquaError(RuntimeError):
    pass


# ---- device-side loops (include/aquasvm.h): the scalar programs of a recorded `while`
class AqsOp(C.Structure):
    _fields_ = [("code", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32),
                ("imm", C.c_double)]


class AqsHeader(C.Structure):
    _fields_ = [("iters", C.c_uint32), ("snaps", C.c_uint32), ("error", C.c_uint32),
                ("cond", C.c_uint32)]


(AQS_IMM, AQS_LOAD, AQS_STORE, AQS_ADD, AQS_SUB, AQS_MUL, AQS_DIV, AQS_MOD, AQS_POW, AQS_NEG, AQS_NOT,
 AQS_LT, AQS_GT, AQS_LE, AQS_GE, AQS_EQ, AQS_NE, AQS_AND, AQS_OR, AQS_SELECT, AQS_CALL, AQS_FOLD,
 AQS_SNAP, AQS_ASSERT, AQS_SETCOND, AQS_RECOND) = range(26)


def aqs_program(ops):
    """[(code, a, b, c, imm), ...] (missing fields = 0; kinds as 'f' / 'u' / 'i') -> AqsOp array."""
    arr = (AqsOp * len(ops))()
    for k, op in enumerate(ops):
        op = tuple(op) + (0,) * (5 - len(op))
        arr[k].code, arr[k].a = op[0], op[1]
        arr[k].b = ord(op[2]) if isinstance(op[2], str) else op[2]
        arr[k].c, arr[k].imm = op[3], float(op[4])
    return arr


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(_LIBPATH)
    L.aqc_last_error.restype = C.c_char_p
    L.aqc_last_error.argtypes = [C.c_void_p]
    L.aqc_get_stream.restype = C.c_void_p
    L.aqc_get_stream.argtypes = [C.c_void_p]
    L.aqc_launch_count.restype = C.c_uint64
    L.aqc_launch_count.argtypes = [C.c_void_p]
    L.aqc_define_round6.restype = C.c_float
    L.aqc_define_round6.argtypes = [C.c_float]
    L.aqc_kernel_name.restype = C.c_char_p
    L.aqc_kernel_args.restype = C.POINTER(ArgInfo)
    L.aqc_kernel_nargs.argtypes = [C.c_int]
    L.aqc_kernel_args.argtypes = [C.c_int]
    L.aqc_kernel_name.argtypes = [C.c_int]
    L.aqc_set_define.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.aqc_kernel_lookup.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.aqc_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.aqc_ctx_destroy.argtypes = [C.c_void_p]
    L.aqc_pairs_cache_enable.argtypes = [C.c_void_p, C.c_int]
    L.aqc_pairs_cache_invalidate.argtypes = [C.c_void_p]
    L.aqc_pairs_cache_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]
    L.aqc_pairs_cache_stats_remote.argtypes = L.aqc_pairs_cache_stats.argtypes
    L.aqc_script_compile.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_char_p),
                                     C.c_int]
    L.aqc_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.aqc_free.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.aqc_host_free.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.aqc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.aqc_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.aqc_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.aqc_set_defs.argtypes = [C.c_void_p, C.POINTER(Defs)]
    L.aqc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_sync.argtypes = [C.c_void_p]
    L.aqc_device_sm_count.argtypes = [C.c_void_p]
    L.aqc_linklist_build.argtypes = [
        C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int,
        C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p,
        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]
    L.aqc_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                 C.c_void_p]
    L.aqc_scatter_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t)]
    L.aqc_reduce.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                             C.c_void_p]
    L.aqc_launch.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_int]
    for n in ("aqc_event_create",):
        getattr(L, n).argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    for n in ("aqc_event_destroy", "aqc_event_record", "aqc_event_sync"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_fused_lookup.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]
    L.aqc_launch_fused.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int]
    L.aqc_comm_unique_id.argtypes = [C.c_void_p]
    L.aqc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.aqc_comm_destroy.argtypes = [C.c_void_p]
    L.aqc_comm_rank.argtypes = [C.c_void_p]
    L.aqc_comm_size.argtypes = [C.c_void_p]
    L.aqc_mpi_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint32)]
    L.aqc_mpi_sync_plan.argtypes = [C.c_void_p]
    L.aqc_mpi_sync_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint32),
                                  C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.aqc_mpi_sync_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.aqc_fp32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.aqc_watch_create.argtypes = [C.c_void_p]
    L.aqc_watch_dirty.argtypes = [C.c_void_p, C.c_int]
    L.aqc_watch_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.aqc_allreduce.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.aqc_allreduce_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.aqc_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.aqc_lane_select.argtypes = [C.c_void_p, C.c_int]
    L.aqc_lane_event.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.aqc_lane_wait.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_loop_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.aqc_loop_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_loop_table.argtypes = [C.c_void_p]
    L.aqc_loop_table.restype = C.c_void_p
    L.aqc_loop_begin.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(AqsOp), C.c_int]
    L.aqc_loop_svm.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(AqsOp), C.c_int]
    L.aqc_loop_end.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_loop_abort.argtypes = [C.c_void_p, C.c_void_p]
    L.aqc_loop_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(AqsHeader),
                               C.c_void_p, C.c_void_p]
    L.aqc_loop_start.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.aqc_loop_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double)]
    L.aqc_kernel_dev_scalars.argtypes = [C.c_int]
    L.aqc_kernel_dev_scalars.restype = C.c_uint64
    L.aqc_launch_ex.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_int,
                                C.POINTER(C.c_void_p)]
    _lib = L
    return L


def kernel_table():
    """[(name, [(argname, type, kind), ...]), ...] of the registry (no GPU needed)."""
    L = lib()
    out = []
    for k in range(L.aqc_kernel_count()):
        n = L.aqc_kernel_nargs(k)
        a = L.aqc_kernel_args(k)
        out.append((L.aqc_kernel_name(k).decode(),
                    [(a[i].name.decode(), a[i].type.decode(), a[i].kind) for i in range(n)]))
    return out


def define_round6(v):
    return float(lib().aqc_define_round6(C.c_float(v)))


class DevArray:
    """A device buffer owned through aqc_alloc (ArrayVariable, Variable.cpp:1739-1822)."""

    def __init__(self, ctx, shape, dtype):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        ctx._chk(lib().aqc_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        self._owned = True

    @property
    def elem_bytes(self):
        n = self.shape[0] if self.shape else 1
        return self.nbytes // max(n, 1)

    def set(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype).reshape(self.shape)
        self.ctx._chk(lib().aqc_memcpy_h2d(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes, 1))
        return self

    def get(self):
        out = np.empty(self.shape, self.dtype)
        self.ctx._chk(lib().aqc_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes, 1))
        return out

    def free(self):
        if self._owned and self.ptr:
            lib().aqc_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _scalar_bytes(value, typ, dims):
    """Pack a python scalar/sequence according to the reference type string."""
    t = typ.replace("unsigned int", "uint").strip()
    if t in ("float",):
        return np.array([value], np.float32).tobytes()
    if t in ("uint", "usize", "size_t"):
        return np.array([value], np.uint32).tobytes()
    if t == "int":
        return np.array([value], np.int32).tobytes()
    if t == "vec":
        v = np.zeros(4 if dims == 3 else 2, np.float32)
        a = np.asarray(value, np.float32).ravel()
        v[:min(len(a), len(v))] = a[:len(v)]
        return v.tobytes()
    if t in ("svec4", "uivec4"):
        return np.asarray(value, np.uint32).reshape(4).tobytes()
    if t == "vec4":
        return np.asarray(value, np.float32).reshape(4).tobytes()
    if t in ("svec2", "uivec2"):
        return np.asarray(value, np.uint32).reshape(2).tobytes()
    raise AquaError("unsupported scalar type '%s'" % typ)


def script_check(path, entry="entry", dims=3, base_path="", defines=()):
    """aqc_script_check: compile a run-time script up to the cubin WITHOUT a device; returns its argument
    list as text, raises AquaError with the compiler's message."""
    L = lib()
    L.aqc_script_check.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_int,
                                   C.c_char_p, C.c_size_t]
    arr = (C.c_char_p * max(len(defines), 1))(*[d.encode() for d in defines])
    log = C.create_string_buffer(4096)
    rc = L.aqc_script_check(path.encode(), entry.encode(), int(dims), base_path.encode(), arr, len(defines), log, 4096)
    if rc < 0:
        raise AquaError(log.value.decode())
    return log.value.decode()


class Context:
    """One CUDA device + stream (CalcServer::setupOpenCL, CalcServer.cpp:858-886)."""

    def __init__(self, device=0, dims=3, h=None):
        self.h = C.c_void_p()
        rc = lib().aqc_ctx_create(int(device), C.byref(self.h))
        if rc:
            raise AquaError("aqc_ctx_create failed: %s" % lib().aqc_last_error(None).decode())
        self.dims = dims
        self.defs = None
        if h is not None:
            self.set_defs_from_h(dims, h)

    @classmethod
    def borrow(cls, handle, dims):
        """Wrap an aqc_ctx* owned by someone else (the C++ host): never destroyed here."""
        self = cls.__new__(cls)
        self.h = C.c_void_p(handle)
        self.dims = dims
        self.defs = None
        self._borrowed = True
        return self

    def wrap(self, devptr, shape, dtype):
        """A DevArray view over device memory owned elsewhere."""
        a = DevArray.__new__(DevArray)
        a.ctx, a.shape, a.dtype = self, tuple(shape), np.dtype(dtype)
        a.nbytes = int(np.prod(a.shape)) * a.dtype.itemsize
        a.ptr, a._owned = devptr, False
        return a

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            lib().aqc_ctx_destroy(self.h)
        self.h = None

    def _chk(self, rc):
        if rc:
            raise AquaError("libaquacuda error %d: %s" % (rc, lib().aqc_last_error(self.h).decode()))

    # -- definitions (basic.xml:119-123 evaluated as CalcServer.cpp:245-257 does)
    def set_defs_from_h(self, dims, h):
        h32 = float(np.float32(h))
        d = Defs()
        d.dims = dims
        d.H = define_round6(h32)
        d.CONW = define_round6(float(np.float32(1.0 / (h32 ** dims))))
        d.CONF = define_round6(float(np.float32(1.0 / (h32 ** (dims + 2)))))
        d.SUPPORT = 2.0
        d.DIMS = define_round6(float(dims))
        self.set_defs(d)

    def set_defs(self, d):
        self.defs = d
        self.dims = d.dims
        self._chk(lib().aqc_set_defs(self.h, C.byref(d)))

    # -- memory
    def array(self, host):
        host = np.ascontiguousarray(host)
        return DevArray(self, host.shape, host.dtype).set(host)

    def empty(self, shape, dtype):
        return DevArray(self, shape, dtype)

    def zeros(self, shape, dtype):
        a = DevArray(self, shape, dtype)
        z = np.zeros(1, np.uint32)
        self._chk(lib().aqc_fill(self.h, a.ptr, a.nbytes // 4, 4, z.ctypes.data))
        return a

    def fill(self, arr, value_bytes):
        eb = len(value_bytes)
        buf = C.create_string_buffer(value_bytes, eb)
        self._chk(lib().aqc_fill(self.h, arr.ptr, arr.nbytes // eb, eb, buf))

    def copy(self, dst, src):
        self._chk(lib().aqc_memcpy_d2d(self.h, dst.ptr, src.ptr, min(dst.nbytes, src.nbytes)))

    def sync(self):
        self._chk(lib().aqc_sync(self.h))

    def launch_count(self):
        return int(lib().aqc_launch_count(self.h))

    # -- pair-mask cache of the neighbour sweeps (include/aquacuda.h)
    def pairs_cache(self, on=True):
        self._chk(lib().aqc_pairs_cache_enable(self.h, 1 if on else 0))

    def pairs_cache_invalidate(self):
        self._chk(lib().aqc_pairs_cache_invalidate(self.h))

    def pairs_cache_stats(self, remote=False):
        b, h, n = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        fn = lib().aqc_pairs_cache_stats_remote if remote else lib().aqc_pairs_cache_stats
        self._chk(fn(self.h, C.byref(b), C.byref(h), C.byref(n)))
        return dict(builds=b.value, hits=h.value, bytes=n.value)

    def sm_count(self):
        return int(lib().aqc_device_sm_count(self.h))

    def fp32_peak(self):
        """Measured FP32 TFLOP/s of the device: (scalar FFMA, packed FFMA2)."""
        a, b = C.c_double(0), C.c_double(0)
        self._chk(lib().aqc_fp32_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- multi-device (include/aquacuda.h): one process per GPU, NCCL
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        if lib().aqc_comm_unique_id(buf):
            raise AquaError("aqc_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, rank, size, unique_id):
        self._chk(lib().aqc_comm_init(self.h, int(rank), int(size), unique_id))

    def mpi_sync_plan(self):
        return int(lib().aqc_mpi_sync_plan(self.h))

    def mpi_sync(self, mask, fields, procs=None, plan=-1, deps=()):
        """MPISync::_execute on DevArrays; returns the number of elements received."""
        nf = len(fields)
        ptrs = (C.c_void_p * max(nf, 1))(*[f.ptr for f in fields])
        eb = (C.c_size_t * max(nf, 1))(*[f.elem_bytes for f in fields])
        pr = (C.c_uint * max(len(procs or ()), 1))(*(procs or ()))
        dp = (C.c_void_p * max(len(deps), 1))(*[d.ptr for d in deps])
        db = (C.c_size_t * max(len(deps), 1))(*[d.nbytes for d in deps])
        nrecv = C.c_uint32(0)
        self._chk(lib().aqc_mpi_sync_ex(self.h, int(plan), mask.ptr, mask.shape[0], nf, ptrs, eb,
                                        len(procs or ()), pr if procs else None, C.byref(nrecv),
                                        len(deps), dp, db))
        return int(nrecv.value)

    def mpi_sync_stats(self, plan):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._chk(lib().aqc_mpi_sync_stats(self.h, int(plan), C.byref(a), C.byref(b)))
        return dict(full=a.value, reused=b.value)

    def allreduce(self, op, typ, arr, count):
        self._chk(lib().aqc_allreduce(self.h, op, typ, arr.ptr, count))

    # -- tools
    def linklist(self, r, support, h, icell, ihoc, perm, inv_perm, rmin=None, rmax=None,
                 recompute=True):
        """LinkList tool. ihoc is a DevArray that may be replaced (grown); returns
        (rmin, rmax, ncells, ihoc)."""
        N = r.shape[0]
        fmin = (C.c_float * 4)(*(list(rmin) + [0.0] * 4)[:4]) if rmin is not None else (C.c_float * 4)()
        fmax = (C.c_float * 4)(*(list(rmax) + [0.0] * 4)[:4]) if rmax is not None else (C.c_float * 4)()
        nc = (C.c_uint32 * 4)()
        p = C.c_void_p(ihoc.ptr if ihoc is not None else None)
        cap = C.c_size_t(ihoc.shape[0] if ihoc is not None else 0)
        self._chk(lib().aqc_linklist_build(self.h, r.ptr, N, self.dims, float(support), float(h),
                                           1 if recompute else 0, fmin, fmax, nc, icell.ptr,
                                           C.byref(p), C.byref(cap), perm.ptr, inv_perm.ptr))
        if ihoc is None or p.value != ihoc.ptr:
            if ihoc is not None:
                ihoc.ptr = None  # freed by the library
            new = DevArray.__new__(DevArray)
            new.ctx, new.shape, new.dtype = self, (int(cap.value),), np.dtype(np.uint32)
            new.nbytes, new.ptr, new._owned = int(cap.value) * 4, p.value, True
            ihoc = new
        vs = 4 if self.dims == 3 else 2
        return (np.array(fmin[:vs], np.float32), np.array(fmax[:vs], np.float32),
                np.array(nc[:], np.uint32), ihoc)

    def radix_sort(self, keys, key_max=0, perm=None, inv_perm=None):
        self._chk(lib().aqc_radix_sort(self.h, keys.ptr, keys.shape[0], int(key_max),
                                       perm.ptr if perm is not None else None,
                                       inv_perm.ptr if inv_perm is not None else None))

    def scatter_fields(self, idx, pairs):
        """dst[idx[i]] = src[i] for every (src, dst) pair."""
        n = len(pairs)
        src = (C.c_void_p * n)(*[s.ptr for s, _ in pairs])
        dst = (C.c_void_p * n)(*[d.ptr for _, d in pairs])
        eb = (C.c_size_t * n)(*[s.elem_bytes for s, _ in pairs])
        self._chk(lib().aqc_scatter_fields(self.h, idx.ptr, idx.shape[0], n, src, dst, eb))

    def reduce(self, op, arr, out_dev=None, host=True):
        dt = arr.dtype
        ncomp = arr.shape[1] if len(arr.shape) > 1 else 1
        if dt == np.float32:
            typ = {1: T_F32, 2: T_VEC2, 4: T_VEC4}[ncomp]
        elif dt == np.uint32:
            typ = T_U32
        elif dt == np.int32:
            typ = T_I32
        else:
            raise AquaError("reduce: unsupported dtype %s" % dt)
        out = np.zeros(ncomp, dt)
        self._chk(lib().aqc_reduce(self.h, op, typ, arr.ptr, arr.shape[0],
                                   out_dev.ptr if out_dev is not None else None,
                                   out.ctypes.data if host else None))
        return out if ncomp > 1 else out[0]

    def lookup(self, script, entry="entry"):
        kid = lib().aqc_kernel_lookup(script.encode(), entry.encode(), self.dims)
        if kid < 0:
            raise AquaError("kernel %s::%s is not in the registry" % (script, entry))
        return kid

    def script_compile(self, path, entry="entry", base_path="", defines=()):
        """A script that is not in the registry, compiled at run time (aqc_script_compile): the kernel
        id, usable with launch(kid=...)."""
        arr = (C.c_char_p * max(len(defines), 1))(*[d.encode() for d in defines])
        kid = lib().aqc_script_compile(self.h, path.encode(), entry.encode(), self.dims, base_path.encode(), arr,
                                       len(defines))
        if kid < 0:
            raise AquaError(lib().aqc_last_error(self.h).decode())
        return kid

    def launch(self, script, entry, variables, n=None, kid=None, dev_scalars=None):
        """Kernel tool: bind arguments by NAME from `variables` (dict name ->
        DevArray | python scalar), like Kernel.cpp:497-556.  dev_scalars: {name: device address}
        of scalars the kernel reads when it runs (aqc_launch_ex)."""
        L = lib()
        if kid is None:
            kid = self.lookup(script, entry)
        na = L.aqc_kernel_nargs(kid)
        info = L.aqc_kernel_args(kid)
        argv = (C.c_void_p * na)()
        keep = []
        for k in range(na):
            name = info[k].name.decode()
            if name not in variables:
                raise AquaError("kernel %s::%s needs variable '%s'" % (script, entry, name))
            v = variables[name]
            if info[k].kind == ARG_SCALAR:
                b = C.create_string_buffer(_scalar_bytes(v, info[k].type.decode(), self.dims))
                keep.append(b)
                argv[k] = C.cast(b, C.c_void_p)
            else:
                argv[k] = v.ptr
        if n is None:
            n = int(variables["N"])
        if dev_scalars:
            dv = (C.c_void_p * na)()
            for k in range(na):
                dv[k] = dev_scalars.get(info[k].name.decode())
            self._chk(L.aqc_launch_ex(self.h, kid, int(n), argv, na, dv))
            return
        self._chk(L.aqc_launch(self.h, kid, int(n), argv, na))

    def loop(self, table_bytes, hist_rows=0, max_ops=1024):
        return DeviceLoop(self, table_bytes, hist_rows, max_ops)

    def launch_fused(self, members, variables):
        """Fused launch of [(script, entry), ...] (pipeline order), arguments by name."""
        L = lib()
        ids = [self.lookup(s_, e_) for s_, e_ in members]
        arr = (C.c_int * len(ids))(*ids)
        fid = L.aqc_fused_lookup(arr, len(ids), self.dims)
        if fid < 0:
            raise AquaError("no fused kernel for %s" % (members,))
        argv, keep = [], []
        for kid in ids:
            na = L.aqc_kernel_nargs(kid)
            info = L.aqc_kernel_args(kid)
            for k in range(na):
                v = variables[info[k].name.decode()]
                if info[k].kind == ARG_SCALAR:
                    b = C.create_string_buffer(_scalar_bytes(v, info[k].type.decode(), self.dims))
                    keep.append(b)
                    argv.append(C.cast(b, C.c_void_p))
                else:
                    argv.append(C.c_void_p(v.ptr))
        a = (C.c_void_p * len(argv))(*argv)
        self._chk(L.aqc_launch_fused(self.h, fid, a, len(argv)))

    # -- events
    def event(self):
        e = C.c_void_p()
        self._chk(lib().aqc_event_create(self.h, C.byref(e)))
        return e

    def record(self, e):
        self._chk(lib().aqc_event_record(self.h, e))

    def elapsed_ms(self, a, b):
        self._chk(lib().aqc_event_sync(self.h, b))
        ms = C.c_float()
        self._chk(lib().aqc_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)


class DeviceLoop:
    """aqc_loop_*: a loop body recorded as a CUDA graph WHILE node (include/aquacuda.h)."""

    def __init__(self, ctx, table_bytes, hist_rows=0, max_ops=1024):
        self.ctx, self.table_bytes, self.hist_rows = ctx, table_bytes, hist_rows
        self.h = C.c_void_p()
        ctx._chk(lib().aqc_loop_create(ctx.h, table_bytes, hist_rows, max_ops, C.byref(self.h)))

    def table(self):
        return lib().aqc_loop_table(self.h)

    def begin(self, entry):
        p = aqs_program(entry)
        self.ctx._chk(lib().aqc_loop_begin(self.ctx.h, self.h, p, len(p)))

    def svm(self, ops):
        p = aqs_program(ops)
        self.ctx._chk(lib().aqc_loop_svm(self.ctx.h, self.h, p, len(p)))

    def end(self):
        self.ctx._chk(lib().aqc_loop_end(self.ctx.h, self.h))

    def abort(self):
        self.ctx._chk(lib().aqc_loop_abort(self.ctx.h, self.h))

    def start(self, table, max_iters=1000):
        """Upload the table: programs queued with svm() before begin() now run at once."""
        tab = np.frombuffer(bytes(table), np.uint8).copy()
        assert tab.nbytes == self.table_bytes
        self.ctx._chk(lib().aqc_loop_start(self.ctx.h, self.h, tab.ctypes.data, max_iters))

    def run(self, table=None, max_iters=1000):
        """-> (header, table bytes as uint8 array, history rows [(tool id, table bytes), ...]).
        table=None: the table start() uploaded, as the programs run since left it."""
        if table is None:
            tab = None
        else:
            tab = np.frombuffer(bytes(table), np.uint8).copy()
            assert tab.nbytes == self.table_bytes
        hdr = AqsHeader()
        out = np.zeros(self.table_bytes, np.uint8)
        hist = np.zeros(max(1, self.hist_rows) * (16 + self.table_bytes), np.uint8)
        self.ctx._chk(lib().aqc_loop_run(self.ctx.h, self.h, tab.ctypes.data if tab is not None else None,
                                         max_iters, C.byref(hdr),
                                         out.ctypes.data, hist.ctypes.data if self.hist_rows else None))
        rows = []
        for k in range(min(hdr.snaps, self.hist_rows)):
            row = hist[k * (16 + self.table_bytes):(k + 1) * (16 + self.table_bytes)]
            rows.append((int(row[:4].view(np.int32)[0]), row[16:].copy()))
        return hdr, out, rows

    def stats(self):
        n, a, b = C.c_int(0), C.c_double(0), C.c_double(0)
        lib().aqc_loop_stats(self.h, C.byref(n), C.byref(a), C.byref(b))
        return n.value, a.value, b.value

    def close(self):
        if self.h:
            lib().aqc_loop_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p()
