"""The `python` tool (aquagpusph/CalcServer/Python.cpp:72-205, 295-395): a script's `main()` is
called once per execution and talks to the simulation through a module called `aquagpusph`
with `get(varname, offset=0, n=0)`, `set(varname, value, offset=0, n=0)` and
`log(log_level, message)`; `main()` must return a bool and False stops the simulation.

The reference embeds CPython in its C++ host.  Here the C++ host calls back into the process
that drives it (aqh_set_script_runner, include/aquahost.h) and this module is that process'
side: it provides the `aquagpusph` module, imports each script once (its folder is appended to
sys.path, as Python.cpp:353-368 does) and runs its main().  The same class serves the CPU oracle
interpreter (oracle/interp.py), with a different backend behind get / set.

A backend provides
    py_get(name, offset, n) -> python int / float (scalars), 1-D numpy array (vec scalars),
                               numpy array (array variables; rows [offset, offset + n))
    py_set(name, value, offset, n)
"""
import importlib.util
import os
import sys
import types

import numpy as np


class ScriptError(RuntimeError):
    pass


_CURRENT = []          # stack of backends: get/set of the module act on the innermost one


def _backend():
    if not _CURRENT:
        raise RuntimeError("aquagpusph.get/set called outside a python tool")
    return _CURRENT[-1]


def _module():
    m = sys.modules.get("aquagpusph")
    if m is not None and getattr(m, "__aqua_b200__", False):
        return m
    m = types.ModuleType("aquagpusph")
    m.__aqua_b200__ = True
    m.__doc__ = "get / set / log of the running simulation (Python.cpp:72-176)"

    def get(varname, offset=0, n=0):
        return _backend().py_get(varname, int(offset), int(n))

    def set(varname, value, offset=0, n=0):   # noqa: A001 (the reference's name)
        _backend().py_set(varname, value, int(offset), int(n))

    def log(log_level, message):
        if int(log_level) >= 2:
            sys.stderr.write(str(message))

    m.get, m.set, m.log = get, set, log
    sys.modules["aquagpusph"] = m
    return m


class ScriptRunner:
    """Loads scripts once per path and calls their main() against `backend`."""

    def __init__(self, backend, base_dir=None, roots=()):
        self.backend = backend
        self.base_dir = base_dir or os.getcwd()
        self.roots = list(roots)     # where "resources/..." paths of the presets are looked for
        self.funcs = {}

    def resolve(self, path):
        if os.path.isabs(path):
            return path
        for d in [self.base_dir] + self.roots:
            full = os.path.normpath(os.path.join(d, path))
            if os.path.exists(full) or os.path.exists(full + ".py"):
                return full
        return os.path.normpath(os.path.join(self.base_dir, path))

    def load(self, path):
        _module()
        full = self.resolve(path)
        if not os.path.exists(full) and os.path.exists(full + ".py"):
            full += ".py"       # the reference imports by module name (Python.cpp:352-368)
        if not os.path.exists(full):
            raise ScriptError("Python module \"%s\" cannot be imported" % path)
        folder = os.path.dirname(full)
        if folder not in sys.path:
            sys.path.append(folder)
        name = "aqua_script_%d_%s" % (id(self), os.path.splitext(os.path.basename(full))[0])
        spec = importlib.util.spec_from_file_location(name, full)
        mod = importlib.util.module_from_spec(spec)
        cwd = os.getcwd()
        _CURRENT.append(self.backend)
        try:
            os.chdir(self.base_dir)     # scripts open their data files relative to the case
            spec.loader.exec_module(mod)
        finally:
            os.chdir(cwd)
            _CURRENT.pop()
        fn = getattr(mod, "main", None)
        if not callable(fn):
            raise ScriptError("main() function cannot be found in \"%s\"" % path)
        self.funcs[path] = fn
        return fn

    def run(self, path):
        fn = self.funcs.get(path) or self.load(path)
        cwd = os.getcwd()
        _CURRENT.append(self.backend)
        try:
            os.chdir(self.base_dir)
            res = fn()
        finally:
            os.chdir(cwd)
            _CURRENT.pop()
        if not isinstance(res, (bool, np.bool_)):
            raise ScriptError("main() function returned non boolean variable")
        if not res:
            raise ScriptError("Python invoked simulation stop")


def narrow(value, dtype, ncomp):
    """What Variable::setFromPythonObject accepts: a python number for plain scalars, a 1-D array
    of exactly ncomp components for vectors (Variable.cpp:259-520)."""
    if ncomp == 1:
        if isinstance(value, np.ndarray):
            if value.size != 1:
                raise ValueError("a scalar variable expected a number")
            value = value.reshape(-1)[0]
        return np.dtype(dtype).type(value)
    a = np.asarray(value)
    if a.ndim != 1 or a.shape[0] != ncomp:
        raise ValueError("expected a 1-D array of %d components" % ncomp)
    return a.astype(dtype)
