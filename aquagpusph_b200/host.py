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
    "aqh_fused_groups", "aqh_save", "aqh_wait_savers", "aqh_checkpoint_file", "aqh_n_savers", "aqh_saver_file",
    "aqh_eval", "aqh_scalar_get", "aqh_scalar_set", "aqh_array_info", "aqh_array_download",
    "aqh_array_upload", "aqh_array_devptr", "aqh_set_script_runner", "aqh_variable_type",
    "aqh_eval_svm", "aqh_device_loops", "aqh_device_loop_stats", "aqh_loop_host_reason",
    "aqh_device_loop_timing", "aqh_device_loop_branch_tools", "aqh_lane_schedule",
]


class HostError(RuntimeError):
    pass


_SCRIPT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_char_p)
_ACTIVE = []            # Simulations inside step() / run(), innermost last
_SCRIPT_ERROR = [None]  # the exception a script raised, re-raised by step() / run()


def _script_dispatch(user, tool, path):
    try:
        if not _ACTIVE:
            raise HostError("python tool \"%s\" ran outside Simulation.step()/run()" % tool.decode())
        _ACTIVE[-1]._run_script(path.decode())
        return 0
    except BaseException as e:   # noqa: BLE001 -- nothing may propagate into the C++ caller
        _SCRIPT_ERROR[0] = e
        return 1


_SCRIPT_CB = _SCRIPT_FN(_script_dispatch)


def type_info(type_name, dims):
    """(numpy dtype, components) of a reference type string; 32-bit indices (State.cpp:499-502)."""
    import re
    t = type_name.replace("*", "").strip()
    if t == "vec":
        return np.float32, 4 if dims == 3 else 2
    if t == "ivec":
        return np.int32, 4 if dims == 3 else 2
    if t in ("uivec", "svec"):
        return np.uint32, 4 if dims == 3 else 2
    if t == "matrix":
        return np.float32, 16 if dims == 3 else 4
    m = re.match(r"^(uivec|svec|ivec|vec)(\d+)$", t)
    if m:
        return {"uivec": np.uint32, "svec": np.uint32, "ivec": np.int32, "vec": np.float32}[m.group(1)], \
            int(m.group(2))
    if t in ("unsigned int", "uint", "size_t", "usize", "unsigned long", "ulong"):
        return np.uint32, 1
    if t in ("int", "long", "ssize_t"):
        return np.int32, 1
    if t == "float":
        return np.float32, 1
    raise HostError("unsupported variable type \"%s\"" % type_name)


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
    L.aqh_save.argtypes = [C.c_void_p]
    L.aqh_wait_savers.argtypes = [C.c_void_p]
    L.aqh_checkpoint_file.argtypes = [C.c_void_p]
    L.aqh_checkpoint_file.restype = C.c_char_p
    L.aqh_n_savers.argtypes = [C.c_void_p]
    L.aqh_saver_file.argtypes = [C.c_void_p, C.c_int]
    L.aqh_saver_file.restype = C.c_char_p
    L.aqh_launch_count.argtypes = [C.c_void_p]
    L.aqh_launch_count.restype = C.c_uint64
    L.aqh_fused_groups.argtypes = [C.c_void_p]
    L.aqh_fused_groups.restype = C.c_uint
    L.aqh_cuda_ctx.argtypes = [C.c_void_p]
    L.aqh_cuda_ctx.restype = C.c_void_p
    L.aqh_eval.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.aqh_eval_svm.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.aqh_device_loops.argtypes = [C.c_void_p]
    L.aqh_device_loops.restype = C.c_uint
    L.aqh_device_loop_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.aqh_device_loop_timing.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
    L.aqh_device_loop_branch_tools.argtypes = [C.c_void_p]
    L.aqh_device_loop_branch_tools.restype = C.c_uint
    L.aqh_loop_host_reason.argtypes = [C.c_void_p, C.c_int]
    L.aqh_loop_host_reason.restype = C.c_char_p
    L.aqh_scalar_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.aqh_scalar_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.aqh_array_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]
    L.aqh_array_download.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    L.aqh_array_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    L.aqh_array_devptr.argtypes = [C.c_void_p, C.c_char_p]
    L.aqh_array_devptr.restype = C.c_void_p
    L.aqh_variable_type.argtypes = [C.c_void_p, C.c_char_p]
    L.aqh_variable_type.restype = C.c_char_p
    # type="python" tools: the host calls back into this process (include/aquahost.h); registered
    # once, before any aqh_load, and dispatched to the Simulation that is stepping
    L.aqh_set_script_runner.argtypes = [_SCRIPT_FN, C.c_void_p]
    L.aqh_set_script_runner.restype = None
    L.aqh_set_script_runner(_SCRIPT_CB, None)
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


def evaluate_svm(expr, type="float", decls="", dims=3, dtype=np.float32, n=1):
    """The same value through the stack programs a recorded `while` runs on the device
    (SvmCompiler + aqs_run of include/aquasvm.h, here on the host)."""
    out = np.zeros(n, dtype)
    _chk(lib().aqh_eval_svm(dims, decls.encode(), type.encode(), expr.encode(), out.ctypes.data,
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
                 parse_only=False, script_dir=None, script_roots=()):
        self.h = C.c_void_p()
        self.dims = dims
        # type="python" tools: relative script paths and the scripts' data files live next to the
        # XML unless script_dir says otherwise; script_roots = folders that hold the presets' Scripts/
        self.script_dir = script_dir or os.path.dirname(os.path.abspath(xml_path))
        self.script_roots = tuple(script_roots) + ((root,) if root else ())
        self._scripts = None
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

    def _stepping(self, fn, *args):
        _ACTIVE.append(self)
        _SCRIPT_ERROR[0] = None
        try:
            rc = fn(self.h, *args)
        finally:
            _ACTIVE.pop()
        if rc and _SCRIPT_ERROR[0] is not None:
            e, _SCRIPT_ERROR[0] = _SCRIPT_ERROR[0], None
            raise HostError("%s (%s: %s)" % (lib().aqh_last_error().decode(), type(e).__name__, e)) from e
        _chk(rc)

    def step(self, n=1):
        self._stepping(lib().aqh_step, int(n))

    def run(self):
        self._stepping(lib().aqh_run)

    def save(self, wait=False):
        """FileManager::save: start the <Save> files of every set (asynchronously) and rewrite the
        AQUAgpusph.save.N.xml state file; returns its path."""
        _chk(lib().aqh_save(self.h))
        if wait:
            self.wait_savers()
        return lib().aqh_checkpoint_file(self.h).decode()

    def wait_savers(self):
        _chk(lib().aqh_wait_savers(self.h))

    def saver_files(self):
        L = lib()
        return [L.aqh_saver_file(self.h, i).decode() for i in range(L.aqh_n_savers(self.h))]

    # type="python" tools (aquagpusph_b200/pytool.py): get / set of the `aquagpusph` module
    def _run_script(self, path):
        if self._scripts is None:
            from . import pytool
            self._scripts = pytool.ScriptRunner(self, self.script_dir, self.script_roots)
        self._scripts.run(path)

    def variable_type(self, name):
        t = lib().aqh_variable_type(self.h, name.encode())
        return t.decode() if t else None

    def py_get(self, name, offset=0, n=0):
        t = self.variable_type(name)
        if t is None:
            raise ValueError('Variable "%s" has not been declared' % name)
        dt, nc = type_info(t, self.dims)
        if "*" in t:
            a = self.download(name, dt)
            return a[offset:offset + n] if n else a[offset:]
        v = self.scalar(name, dt, nc)
        if nc > 1:
            return v
        return float(v) if dt == np.float32 else int(v)

    def py_set(self, name, value, offset=0, n=0):
        from . import pytool
        t = self.variable_type(name)
        if t is None:
            raise ValueError('Variable "%s" has not been declared' % name)
        dt, nc = type_info(t, self.dims)
        if "*" in t:
            a = np.ascontiguousarray(value, dt)
            full = self.download(name, dt) if (offset or a.shape[0] != self.array_info(name)[0]) else a
            if full is not a:
                full[offset:offset + a.shape[0]] = a
            self.upload(name, full)
            return
        v = np.atleast_1d(pytool.narrow(value, dt, nc))
        # exact: repr of the float32 value read as a double narrows back to the same float32
        self.set_scalar(name, ", ".join(repr(float(x)) if dt == np.float32 else str(int(x)) for x in v))

    def sync(self):
        _chk(lib().aqh_sync(self.h))

    def launch_count(self):
        return int(lib().aqh_launch_count(self.h))

    def fused_groups(self):
        return int(lib().aqh_fused_groups(self.h))

    def cuda_ctx(self):
        return lib().aqh_cuda_ctx(self.h)

    def device_loops(self):
        """`while` loops that run as a CUDA graph from their second pass on."""
        return int(lib().aqh_device_loops(self.h))

    def device_loop_stats(self):
        """(times a loop ran on the device, passes made there)."""
        runs, iters = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().aqh_device_loop_stats(self.h, C.byref(runs), C.byref(iters)))
        return int(runs.value), int(iters.value)

    def device_loop_branch_tools(self):
        """Tools of the device loops that run on the second stream."""
        return int(lib().aqh_device_loop_branch_tools(self.h))

    def device_loop_timing(self):
        """{graph nodes of the recorded bodies, host ms of the last recording / instantiation}."""
        n, r, i = C.c_int(0), C.c_double(0), C.c_double(0)
        _chk(lib().aqh_device_loop_timing(self.h, C.byref(n), C.byref(r), C.byref(i)))
        return {"body_nodes": n.value, "record_ms": r.value, "instantiate_ms": i.value}

    def loop_host_reason(self, i):
        """Why the `while` at tool index i stays on the host ('' when it runs on the device)."""
        r = lib().aqh_loop_host_reason(self.h, int(i))
        return None if r is None else r.decode()

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
