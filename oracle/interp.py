"""CPU ORACLE host (test infrastructure, NOT product code).

An independent, pure-Python re-implementation of the reference's per-step
pipeline semantics (CalcServer::update, CalcServer.cpp:592-621; tool types of
aquagpusph/CalcServer/*.cpp) that interprets a RESOLVED problem XML
(aquagpusph_b200/cases_xml/*.xml instantiated by casegen.instantiate) and runs
every device tool through the C oracle (oracle/liboracle.so).  It shares no code
with the C++ host or the CUDA library, so agreement after N steps checks the
XML front-end, the scheduler semantics and the kernels at once.  It is also the
"port" CPU baseline that bench.py times.
"""
import math
import re
import xml.etree.ElementTree as ET

import os

import numpy as np

from . import oracle as O

_EXT = ["_x", "_y", "_z", "_w"]


# ---- expressions (Tokenizer semantics: double precision, then narrowed) -------
class _Expr:
    TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)[fF]?|"
                     r"([A-Za-z_]\w*)|(<=|>=|==|!=|&&|\|\||[-+*/^%()<>?:,!]))")

    def __init__(self, s, env):
        self.toks = []
        pos = 0
        s = s.strip()
        while pos < len(s):
            m = self.TOK.match(s, pos)
            if not m:
                raise ValueError("cannot tokenize %r at %d" % (s, pos))
            self.toks.append(m.groups())
            pos = m.end()
        self.i = 0
        self.env = env

    def peek(self):
        return self.toks[self.i][2] if self.i < len(self.toks) else None

    def eat(self, op):
        if self.peek() == op:
            self.i += 1
            return True
        return False

    def parse(self):
        v = self.ternary()
        if self.i != len(self.toks):
            raise ValueError("trailing tokens")
        return v

    def ternary(self):
        c = self.lor()
        if self.eat("?"):
            a = self.ternary()
            assert self.eat(":")
            b = self.ternary()
            return a if c != 0 else b
        return c

    def lor(self):
        a = self.land()
        while self.eat("||"):
            b = self.land()
            a = 1.0 if (a != 0 or b != 0) else 0.0
        return a

    def land(self):
        a = self.cmp()
        while self.eat("&&"):
            b = self.cmp()
            a = 1.0 if (a != 0 and b != 0) else 0.0
        return a

    def cmp(self):
        a = self.add()
        while self.peek() in ("<=", ">=", "==", "!=", "<", ">"):
            op = self.peek()
            self.i += 1
            b = self.add()
            a = float({"<=": a <= b, ">=": a >= b, "==": a == b, "!=": a != b,
                       "<": a < b, ">": a > b}[op])
        return a

    def add(self):
        a = self.mul()
        while self.peek() in ("+", "-"):
            op = self.peek()
            self.i += 1
            b = self.mul()
            a = a + b if op == "+" else a - b
        return a

    def mul(self):
        a = self.unary()
        while self.peek() in ("*", "/", "%"):
            op = self.peek()
            self.i += 1
            b = self.unary()
            if op == "*":
                a = a * b
            elif op == "/":  # IEEE semantics like the C++ evaluators (x/0 = inf, 0/0 = nan)
                with np.errstate(all="ignore"):
                    a = float(np.float64(a) / np.float64(b))
            else:
                a = math.fmod(a, b)
        return a

    def unary(self):
        if self.eat("-"):
            return -self.unary()
        if self.eat("+"):
            return self.unary()
        if self.eat("!"):
            return 1.0 if self.unary() == 0 else 0.0
        return self.power()

    def power(self):
        b = self.primary()
        if self.eat("^"):
            return math.pow(b, self.unary())
        return b

    def primary(self):
        num, ident, op = self.toks[self.i]
        self.i += 1
        if num is not None:
            return float(num)
        if op == "(":
            v = self.ternary()
            assert self.eat(")")
            return v
        if ident is not None:
            if self.eat("("):
                args = []
                if not self.eat(")"):
                    while True:
                        args.append(self.ternary())
                        if self.eat(","):
                            continue
                        assert self.eat(")")
                        break
                f = {"sqrt": math.sqrt, "abs": abs, "sin": math.sin, "cos": math.cos,
                     "tan": math.tan, "exp": math.exp, "log": math.log, "floor": math.floor,
                     "ceil": math.ceil, "pow": math.pow, "min": min, "max": max}[ident]
                return float(f(*args))
            return float(self.env[ident])
        raise ValueError("unexpected token %r" % op)


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        depth += ch == "("
        depth -= ch == ")"
        if ch in ",;" and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


_LITERALS = {"VEC_ZERO": (0.0, True), "VEC_ONE": (1.0, False), "VEC_ALL_ONE": (1.0, True),
             "VEC_INFINITY": (math.inf, False), "VEC_ALL_INFINITY": (math.inf, True),
             "INFINITY": (math.inf, True), "MAT_ZERO": (0.0, True)}


class LocalTransport:
    """N interpreter ranks inside one process (one Python thread per rank)."""

    def __init__(self, size):
        import threading
        self.size = size
        self._slots = [None] * size
        self._barrier = threading.Barrier(size)

    def allgather(self, rank, obj):
        self._slots[rank] = obj
        self._barrier.wait()
        out = list(self._slots)
        self._barrier.wait()
        return out


class TorchTransport:
    """One interpreter rank per process over torch.distributed (gloo on CPU)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.size = dist.get_world_size()

    def allgather(self, rank, obj):
        out = [None] * self.size
        self.dist.all_gather_object(out, obj)
        return out


class Interpreter:
    def __init__(self, xml_text, dims, rank=0, size=1, transport=None):
        self.dims = dims
        self.rank, self.size, self.transport = rank, size, transport
        if size > 1 and transport is None:
            raise ValueError("a multi-rank interpreter needs a transport")
        self._ref = None
        self.root = ET.fromstring(re.sub(r"<!--.*?-->", "", xml_text, flags=re.S))
        self.sets = []
        for s in self.root.iter("ParticlesSet"):
            self.sets.append((int(s.get("n")), [(c.get("name"), c.get("value")) for c in s.iter("Scalar")]))
        self.N = sum(n for n, _ in self.sets)
        self.env = {"pi": math.pi, "e": math.e, "INFINITY": math.inf}
        self.types = {}
        self.V = {}
        N = self.N
        n_radix = 1
        while n_radix < ((N + 1023) // 1024) * 1024:
            n_radix *= 2
        for name, typ, val in [("mpi_rank", "unsigned int", str(rank)), ("mpi_size", "unsigned int", str(size)),
                               ("dims", "unsigned int", str(dims)), ("t", "float", "0"),
                               ("dt", "float", "0"), ("iter", "unsigned int", "0"),
                               ("frame", "unsigned int", "0"), ("N", "size_t", str(N)),
                               ("n_sets", "unsigned int", str(len(self.sets))),
                               ("n_radix", "size_t", str(n_radix)), ("n_cells", "svec4", "1,1,1,1"),
                               ("support", "float", "2")]:
            self.reg_scalar(name, typ, val)
        for name, typ in [("id", "size_t*"), ("r", "vec*"), ("iset", "unsigned int*"),
                          ("id_sorted", "size_t*"), ("id_unsorted", "size_t*"), ("icell", "size_t*")]:
            self.reg_array(name, typ, "N")
        self.types["ihoc"] = "size_t*"
        self.V["ihoc"] = np.zeros(1, np.uint32)
        for v in self.root.iter("Variable"):
            if "*" in v.get("type"):
                self.reg_array(v.get("name"), v.get("type"), v.get("length"))
            else:
                self.reg_scalar(v.get("name"), v.get("type"), v.get("value") or "")
        # definitions (CalcServer.cpp:240-265)
        self.defs = {}
        for d in self.root.iter("Define"):
            val = d.get("value")
            if val is not None and d.get("evaluate") == "true":
                f = float(np.float32(self.eval(val)))
                val = repr(float(O.lib().aqo_define_round6(O.C.c_float(f))))
            self.defs[d.get("name")] = val
        self.D = O.Defs()
        self.D.dims = dims
        self.D.H = float(self.defs["H"])
        self.D.CONW = float(self.defs["CONW"])
        self.D.CONF = float(self.defs["CONF"])
        self.D.SUPPORT = 2.0
        self.dr_factor = float(str(self.defs.get("__DR_FACTOR__", "0.5f")).rstrip("f"))
        # default particle data (Particles::loadDefault)
        off = 0
        for k, (n, _) in enumerate(self.sets):
            self.V["iset"][off:off + n] = k
            off += n
        for k in ("id", "id_sorted", "id_unsorted"):
            self.V[k][:] = np.arange(N, dtype=np.uint32)
        for k, (n, scalars) in enumerate(self.sets):
            for name, val in scalars:
                self.V[name][k] = np.float32(self.eval(val))
        self.tools = [dict(t.attrib, operation=((t.text or "").strip() or t.attrib.get("operation", "")))
                      for t in self.root.iter("Tool")]
        self.once_done = set()
        self.if_state = {}
        self.steps = 0
        self._scopes()

    # -- variables
    def _base(self, typ):
        t = typ.replace("*", "").strip()
        t = {"size_t": "unsigned int", "uint": "unsigned int", "svec4": "uivec4"}.get(t, t)
        if t == "vec":
            t = "vec4" if self.dims == 3 else "vec2"
        if t == "matrix":
            return np.float32, 16 if self.dims == 3 else 4
        m = re.match(r"(uivec|ivec|vec)(\d)", t)
        if m:
            return {"uivec": np.uint32, "ivec": np.int32, "vec": np.float32}[m.group(1)], int(m.group(2))
        return {"unsigned int": np.uint32, "int": np.int32, "float": np.float32}[t], 1

    def reg_array(self, name, typ, length):
        dt, n = self._base(typ)
        L = int(self.eval(length))
        self.types[name] = typ
        self.V[name] = np.zeros((L, n) if n > 1 else (L,), dt)

    def reg_scalar(self, name, typ, val):
        dt, n = self._base(typ)
        self.types[name] = typ
        self.V[name] = np.zeros(n, dt) if n > 1 else dt(0)
        if val.strip():
            self.set_scalar(name, val)
        else:
            self.publish(name)

    def eval(self, expr):
        return _Expr(expr, self.env).parse()

    def publish(self, name):
        v = self.V[name]
        if np.ndim(v) == 0:
            self.env[name] = float(v)
        else:
            for c in range(len(v)):
                self.env[name + _EXT[c]] = float(v[c])

    def _narrow(self, dt, x):
        if dt == np.float32:
            return np.float32(x)
        return dt(int(x))

    def set_scalar(self, name, expr):
        dt, n = self._base(self.types[name])
        if n == 1:
            self.V[name] = self._narrow(dt, self.eval(expr))
        else:
            parts = _split_top(expr)
            self.V[name] = np.array([self._narrow(dt, self.eval(p)) for p in parts[:n]], dt)
        self.publish(name)

    def element_value(self, typ, value):
        dt, n = self._base(typ)
        v = value.strip()
        m = re.match(r"^\(\((\w+)\)\((.*)\)\)$", v)
        if m:
            v = m.group(2)
        neg = v.startswith("-") and v[1:].strip() in _LITERALS
        key = v[1:].strip() if neg else v
        if key in _LITERALS:
            x, allc = _LITERALS[key]
            out = np.full(n, -x if neg else x, dt)
            if not allc and self.dims == 3 and n == 4:
                out[3] = 0
            return out if n > 1 else out[0]
        parts = _split_top(v)
        if len(parts) == 1:
            return np.full(n, self._narrow(dt, self.eval(v)), dt) if n > 1 else self._narrow(dt, self.eval(v))
        return np.array([self._narrow(dt, self.eval(p)) for p in parts[:n]], dt)

    # -- control flow (Conditional.cpp)
    def _scopes(self):
        self.jump_end, self.jump_back = {}, {}
        stack = []
        for i, t in enumerate(self.tools):
            if t["type"] in ("if", "while"):
                stack.append(i)
            elif t["type"] in ("end", "endif"):
                o = stack.pop()
                self.jump_end[o] = i + 1
                self.jump_back[i] = o

    def ll(self):
        return O.make_ll(self.V["icell"], self.V["ihoc"], self.V["n_cells"], self.N)

    def step(self):
        i = 0
        n = len(self.tools)
        while i < n:
            i = self.run_tool(i)
        self.steps += 1

    def run_tool(self, i):
        t = self.tools[i]
        typ = t["type"]
        if t.get("once") == "true":
            if i in self.once_done:
                return i + 1
            self.once_done.add(i)
        V = self.V
        if typ in ("dummy",) or typ.startswith("report_"):
            return i + 1
        if typ == "kernel":
            self.kernel(t["path"], t.get("entry_point", "entry"))
        elif typ == "copy":
            V[t["out"]][...] = V[t["in"]]
        elif typ == "set":
            V[t["in"]][...] = self.element_value(self.types[t["in"]], t["value"])
        elif typ == "set_scalar":
            self.set_scalar(t["in"], t["value"])
        elif typ == "assert":
            if self.eval(t["condition"]) == 0:
                raise AssertionError("Assertion error in tool %s: %s" % (t["name"], t["condition"]))
        elif typ == "while":
            if self.eval(t["condition"]) == 0:
                return self.jump_end[i]
        elif typ == "if":
            st = self.if_state.get(i, True)
            if st:
                res = self.eval(t["condition"]) != 0
                self.if_state[i] = not res  # flipped for the visit coming back from End
                if not res:
                    self.if_state[i] = True
                    return self.jump_end[i]
            else:
                self.if_state[i] = True
                return self.jump_end[i]
        elif typ in ("end", "endif"):
            return self.jump_back[i]
        elif typ == "reduction":
            self.reduction(t)
        elif typ == "radix-sort":
            # RadixSort.cpp:129-303: keys sorted in place (ascending, stable); perm[sorted] = original
            # index, inv_perm[original] = sorted index (RadixSort.cl.in:313-323)
            keys = V[t["in"]]
            perm = np.argsort(keys, kind="stable").astype(np.uint32)
            keys[...] = keys[perm]
            if t.get("perm"):
                V[t["perm"]][...] = perm
            if t.get("inv_perm"):
                V[t["inv_perm"]][perm] = np.arange(len(perm), dtype=np.uint32)
        elif typ == "link-list":
            self.linklist(t)
        elif typ == "mpi-sync":
            self.mpi_sync(t)
        elif typ == "mpi-allreduce":
            self.mpi_allreduce(t)
        elif typ == "python":
            self.python(t)
        else:
            raise NotImplementedError("oracle interpreter: tool type %s" % typ)
        return i + 1

    def reduction(self, t):
        a = self.V[t["in"]]
        op = re.sub(r"[\s;]", "", t["operation"])
        ident = self.element_value(self.types[t["in"]], t["null"])
        if op in ("c=a+b",):
            if a.ndim == 1:
                r = O.lib().aqo_reduce_sum_tree(O._arg(np.ascontiguousarray(a, np.float32)), a.shape[0], 256) \
                    if a.dtype == np.float32 else a.sum(dtype=np.uint64)
            else:
                r = np.zeros(a.shape[1], np.float32)
                O.call("reduce_sum_vec_tree", np.ascontiguousarray(a), a.shape[0], a.shape[1], 256, r)
            r = r + ident
        elif op in ("c=min(a,b)", "c=(a<b)?a:b", "c=(a>b)?b:a"):
            r = np.minimum(a.min(0), ident)
        elif op in ("c=max(a,b)", "c=(a<b)?b:a", "c=(a>b)?a:b"):
            r = np.maximum(a.max(0), ident)
        else:
            raise NotImplementedError(op)
        dt, n = self._base(self.types[t["out"]])
        self.V[t["out"]] = np.asarray(r, dt) if n > 1 else dt(r)
        self.publish(t["out"])

    def linklist(self, t):
        """LinkList tool with the attribute defaults of State.cpp:1104-1114."""
        V = self.V
        r = V[t.get("in", "r")]
        vmin, vmax = t.get("min", "r_min"), t.get("max", "r_max")
        recompute = (t.get("recompute_grid", "true").lower() != "false")
        rmin = rmax = None
        if recompute:
            rmin = np.zeros(O.vs(self.dims), np.float32)
            rmax = np.zeros(O.vs(self.dims), np.float32)
            O.call("minmax", np.ascontiguousarray(r), r.shape[0], self.dims, rmin, rmax)
            if self.size > 1:
                # ADDITION to the reference (one global grid, see csrc/mpi.cu aqc_comm_minmax)
                allmm = self.transport.allgather(self.rank, (rmin, rmax))
                rmin = np.min([a for a, _ in allmm], axis=0)
                rmax = np.max([b for _, b in allmm], axis=0)
        else:
            rmin, rmax = np.array(V[vmin], np.float32), np.array(V[vmax], np.float32)
        res = O.linklist(r, self.dims, float(V["support"]), float(V["h"]), rmin, rmax, recompute=False)
        if recompute:
            V[vmin], V[vmax] = res["rmin"], res["rmax"]
            self.publish(vmin)
            self.publish(vmax)
        nc = t.get("n_cells", "n_cells")
        V[nc] = res["ncells"]
        self.publish(nc)
        icell, ihoc = t.get("icell", "icell"), t.get("ihoc", "ihoc")
        perm, inv = t.get("perm", "id_unsorted"), t.get("inv_perm", "id_sorted")
        n = r.shape[0]
        for name, val in ((icell, res["icell"]), (perm, res["perm"]), (inv, res["inv_perm"])):
            if V[name].shape[0] == n:
                V[name] = val
            else:
                V[name][:n] = val
        if ihoc == "ihoc" or V[ihoc].shape[0] < res["ihoc"].shape[0]:
            V[ihoc] = res["ihoc"]       # reallocatable (LinkList.cpp:234-271)
        else:
            V[ihoc][:res["ihoc"].shape[0]] = res["ihoc"]

    def mpi_sync(self, t):
        """MPISync::_execute (MPISync.cpp:183-232) on host arrays."""
        if self.size <= 1:
            return
        V = self.V
        mask = V[t["mask"]]
        fields = [f.strip() for f in t["fields"].split(",") if f.strip()]
        procs = [int(self.eval(p)) for p in t.get("processes", "").split(",") if p.strip()]
        if not procs:
            procs = [p for p in range(self.size) if p != self.rank]
        perm = np.argsort(mask, kind="stable")
        smask = mask[perm]
        sends = {}
        for p in procs:
            if p == self.rank:
                continue
            sel = perm[smask == p]          # sorted block bound to p, original order kept
            sends[p] = [V[f][sel].copy() for f in fields]
        got = self.transport.allgather(self.rank, sends)
        mask[...] = self.rank
        off = 0
        for p in sorted(procs):
            if p == self.rank:
                continue
            data = got[p].get(self.rank) if isinstance(got[p], dict) else None
            if not data or not len(data[0]):
                continue
            n = len(data[0])
            for f, a in zip(fields, data):
                V[f][off:off + n] = a
            mask[off:off + n] = p
            off += n

    # -- type="python" (Python.cpp:295-325): the script's main() against this interpreter's variables
    script_dir = None       # where relative script paths and the scripts' data files live
    script_roots = ()       # where the presets' "Scripts/..." paths are looked for

    def python(self, t):
        if self.script_dir is None:
            raise NotImplementedError("oracle interpreter: tool type python needs script_dir")
        if getattr(self, "_scripts", None) is None:
            from aquagpusph_b200 import pytool
            self._scripts = pytool.ScriptRunner(self, self.script_dir, self.script_roots)
        self._scripts.run(t["path"])

    def py_get(self, name, offset=0, n=0):
        if name not in self.V:
            raise ValueError('Variable "%s" has not been declared' % name)
        v = self.V[name]
        if "*" in self.types[name]:
            return np.array(v[offset:offset + n] if n else v[offset:])
        if np.ndim(v) == 0:
            return float(v) if isinstance(v, np.floating) else int(v)
        return np.array(v)

    def py_set(self, name, value, offset=0, n=0):
        if name not in self.V:
            raise ValueError('Variable "%s" has not been declared' % name)
        if "*" in self.types[name]:
            a = np.asarray(value)
            self.V[name][offset:offset + a.shape[0]] = a
            return
        from aquagpusph_b200 import pytool
        dt, nc = self._base(self.types[name])
        self.V[name] = pytool.narrow(value, dt, nc)
        self.publish(name)

    def mpi_allreduce(self, t):
        if self.size <= 1:
            return
        name, op = t["in"], t.get("operation", "min").strip().lower()
        vals = self.transport.allgather(self.rank, np.array(self.V[name]))
        fn = {"min": np.min, "max": np.max, "sum": np.sum, "add": np.sum}[op]
        r = fn(np.array(vals), axis=0)
        dt, n = self._base(self.types[name])
        self.V[name] = np.asarray(r, dt) if n > 1 else dt(r)
        self.publish(name)

    # -- script kernels
    def kernel(self, path, entry):
        rel = path.split("Scripts/")[-1]
        V, N, d = self.V, self.N, self.dims
        D = self.D
        key = (rel, entry)

        def c(name, *args):  # every script kernel is a loop over independent rows
            O.pcall(name, N, *args)

        f32 = lambda k: float(V[k])  # noqa: E731
        if key in (("basic/time_scheme/midpoint.cl", "predictor"), ("basic/time_scheme/euler.cl", "predictor")):
            c("mp_predictor", V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"], V["r_in"], V["u_in"],
              V["dudt_in"], V["rho_in"], V["drhodt_in"], N, d)
        elif key == ("basic/time_scheme/improved_euler.cl", "predictor"):
            c("ie_predictor", V["imove"], V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"], V["r_in"],
              V["u_in"], V["dudt_in"], V["rho_in"], V["drhodt_in"], N, f32("dt"), d)
        elif key == ("basic/time_scheme/improved_euler.cl", "corrector"):
            c("ie_corrector", V["imove"], V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"],
              V["dudt_in"], V["drhodt_in"], N, f32("dt"), d)
        elif key == ("basic/time_scheme/euler.cl", "corrector"):
            c("euler_corrector", V["imove"], V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"], N,
              f32("dt"), d)
        elif key == ("basic/Domain.cl", "entry"):
            c("domain", V["imove"], V["r_in"], V["u_in"], V["dudt_in"], V["m"], N, V["domain_min"],
              V["domain_max"], d)
        elif key == ("basic/Sort.cl", "stage1"):
            for k in ("id", "iset", "imove", "r", "normal", "tangent"):
                V[k] = O.scatter(V[k + "_in"], V["id_sorted"])
        elif key == ("basic/Sort.cl", "stage2"):
            for k in ("rho", "m", "u"):
                V[k] = O.scatter(V[k + "_in"], V["id_sorted"])
            V["dudt_in"] = O.scatter(V["dudt"], V["id_sorted"])
            V["drhodt_in"] = O.scatter(V["drhodt"], V["id_sorted"])
        elif key == ("basic/EOS.cl", "entry"):
            c("eos", V["iset"], V["imove"], V["rho"], V["p"], V["refd"], N, f32("cs"), f32("p0"))
        elif key == ("basic/Binormal.cl", "entry"):
            c("binormal", V["normal"], V["tangent"], V["binormal"], N, d)
        elif key == ("basic/neighs.cl", "entry"):
            c("neighs", self.ll(), V["imove"], V["n_neighs"], int(V["neighs_limit"]), d)
        elif key == ("basic/MLS.cl", "entry"):
            c("mls", D, self.ll(), V["imove"], V["r"], V["rho"], V["m"], V["mls"], int(V["mls_imove"]))
        elif key == ("basic/MLS.cl", "mls_inv"):
            c("mls_inv", V["imove"], V["mls"], N, int(V["mls_imove"]), d)
        elif key == ("cfd/Shepard.cl", "entry"):
            c("shepard", D, self.ll(), 1, V["imove"], V["r"], V["rho"], V["m"], V["shepard"])
        elif key == ("cfd/Interactions.cl", "entry"):
            lap = str(self.defs.get("__LAP_FORMULATION__", "__LAP_MONAGHAN__")).strip()
            lap = str(self.defs.get(lap, lap)).strip()      # (__LAP_MORRIS__ -> 2)
            if lap not in ("1", "2", "__LAP_MONAGHAN__", "__LAP_MORRIS__"):
                raise NotImplementedError("oracle interpreter: __LAP_FORMULATION__=" + lap)
            c("interactions_morris" if lap in ("2", "__LAP_MORRIS__") else "interactions", D, self.ll(), V["imove"],
              V["r"], V["u"], V["rho"], V["m"], V["p"], V["grad_p"], V["lap_u"], V["div_u"])
        elif key == ("cfd/Sensors.cl", "entry"):
            c("sensors", D, self.ll(), V["imove"], V["r"], V["m"], V["u"], V["rho"], V["p"])
        elif key == ("cfd/SensorsRenormalization.cl", "entry"):
            c("sensors_renorm", V["imove"], V["shepard"], V["u"], V["rho"], V["p"], N, d)
        elif key == ("cfd/deltaSPH.cl", "full"):
            c("dsph_full", D, self.ll(), V["imove"], V["r"], V["rho"], V["m"], V["p"], V["lap_p_corr"])
        elif key == ("cfd/deltaSPH.cl", "lapp"):
            c("dsph_lapp", D, self.ll(), V["imove"], V["r"], V["rho"], V["m"], V["p"], V["lap_p"])
        elif key == ("cfd/deltaSPH.cl", "full_mls"):
            c("dsph_full_mls", V["imove"], V["mls"], V["lap_p_corr"], N, d)
        elif key == ("cfd/deltaSPH.cl", "lapp_corr"):
            c("dsph_lapp_corr", D, self.ll(), V["imove"], V["r"], V["rho"], V["m"], V["lap_p_corr"],
              V["lap_p"])
        elif key == ("cfd/deltaSPH.cl", "deltaSPH"):
            c("dsph_apply", V["iset"], V["imove"], V["rho"], V["lap_p"], V["drhodt"], V["refd"],
              V["delta"], N, f32("dt"))
        elif key == ("cfd/Rates.cl", "entry"):
            c("rates", V["iset"], V["imove"], V["rho"], V["grad_p"], V["lap_u"], V["div_u"], V["dudt"],
              V["drhodt"], V["visc_dyn"], N, V["g"], d)
        elif key == ("cfd/TimeStep.cl", "entry"):
            c("timestep", V["imove"], V["u"], V["dt_var"], N, f32("dt"), f32("dt_min"), f32("courant"),
              f32("dt_Ma"), f32("h"), d)
        elif key == ("cfd/Boundary/BIe/Interactions.cl", "entry"):
            c("bie_interactions", D, self.ll(), V["imove"], V["r"], V["normal"], V["u"], V["m"],
              V["grad_w_bi"], V["div_u_bi"])
        elif key == ("cfd/Boundary/BIe/Interactions.cl", "p_boundary"):
            c("bie_p_boundary", D, self.ll(), V["imove"], V["r"], V["m"], V["rho"], V["p"])
        elif key == ("cfd/Boundary/BIe/Rates.cl", "entry"):
            c("bie_rates", V["imove"], V["rho"], V["p"], V["u"], V["grad_w_bi"], V["div_u_bi"],
              V["grad_p"], V["div_u"], N, d)
        elif key == ("cfd/Boundary/BIe/Rates.cl", "force_press"):
            c("bie_force_press", V["imove"], V["r"], V["normal"], V["m"], V["p"], V["force_p"],
              V["moment_p"], V["forces_r"], N, d)
        elif key == ("cfd/Boundary/BIe/ElasticBounce.cl", "entry"):
            self._bie_eb()
        elif key == ("cfd/Boundary/BIe/ElasticBounce.cl", "force_bound"):
            c("bie_force_bound", V["imove"], V["m"], V["dudt_preelastic"], V["dudt_elastic"],
              V["force_elastic"], N, d)
        elif key == ("cfd/Boundary/BIe/PST.cl", "entry"):
            self._bie_pst()
        elif key == ("basic/time_scheme/midpoint.cl", "midpoint"):
            c("mp_midpoint", V["imove"], V["u_in"], V["u"], V["dudt"], V["rho_in"], V["rho"],
              V["drhodt"], N, f32("dt"), d)
        elif key == ("basic/time_scheme/midpoint.cl", "relax"):
            c("mp_relax", V["imove"], V["dudt_in"], V["dudt"], V["drhodt_in"], V["drhodt"], N,
              f32("relax_midpoint"), d)
        elif key == ("basic/time_scheme/midpoint.cl", "residuals"):
            c("mp_residuals", V["imove"], V["m"], V["u"], V["dudt_in"], V["dudt"], V["rho"], V["p"],
              V["drhodt_in"], V["drhodt"], V["residual_midpoint"], N, d)
        elif key == ("basic/time_scheme/midpoint.cl", "corrector"):
            c("mp_corrector", V["imove"], V["r_in"], V["r"], V["u_in"], V["u"], V["dudt"], V["rho_in"],
              V["rho"], V["drhodt"], N, f32("dt"), d)
        elif key == ("cfd/Motions/Transform.cl", "entry"):
            c("motion_transform", V["iset"], V["imove"], V["r"], V["normal"], V["tangent"], N,
              int(V["motion_iset"]), V["motion_r"], V["motion_a"], d)
        elif key == ("cfd/Motions/UnTransform.cl", "entry"):
            c("motion_untransform", V["iset"], V["imove"], V["r"], V["normal"], V["tangent"], N,
              int(V["motion_iset"]), V["motion_r_in"], V["motion_a_in"], d)
        elif key == ("cfd/Motions/Velocity.cl", "entry"):
            c("motion_velocity", V["iset"], V["imove"], V["r"], V["u"], N, int(V["motion_iset"]),
              V["motion_drdt"], V["motion_a"], V["motion_dadt"], d)
        elif key == ("cfd/Motions/Acceleration.cl", "entry"):
            c("motion_acceleration", V["iset"], V["imove"], V["r"], V["dudt"], N, int(V["motion_iset"]),
              V["motion_ddrddt"], V["motion_a"], V["motion_ddaddt"], d)
        elif key == ("cfd/Energy/Energy.cl", "power"):
            c("energy_power", V["energy_dekdt"], V["energy_depdt"], V["energy_decdt"], V["imove"], V["u"],
              V["rho"], V["m"], V["p"], V["dudt"], V["drhodt"], N, V["g"], d)
        elif key == ("cfd/Energy/Energy.cl", "energy"):
            c("energy_energy", V["energy_ek"], V["energy_ep"], V["energy_ec"], V["iset"], V["imove"],
              V["r"], V["u"], V["rho"], V["m"], V["refd"], N, V["g"], f32("cs"), d)
        elif key == ("basic/time_scheme/adam_bashforth.cl", "predictor"):
            c("mp_predictor", V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"], V["r_in"], V["u_in"],
              V["dudt_in"], V["rho_in"], V["drhodt_in"], N, d)
        elif key == ("basic/time_scheme/adam_bashforth.cl", "sort"):
            lv = range(1, 5)
            O.call("ab_sort", O.ptrs(V["dudt_as%d_in" % l] for l in lv), O.ptrs(V["dudt_as%d" % l] for l in lv),
                   O.ptrs(V["drhodt_as%d_in" % l] for l in lv), O.ptrs(V["drhodt_as%d" % l] for l in lv),
                   V["id_sorted"], N, d)
        elif key == ("basic/time_scheme/adam_bashforth.cl", "corrector"):
            lv = range(1, 5)
            steps = int(str(self.defs.get("TSCHEME_ADAMS_BASHFORTH_STEPS", "5u")).rstrip("uU"))
            O.call("ab_corrector", V["imove"], V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"],
                   O.ptrs(V["dudt_as%d" % l] for l in lv), O.ptrs(V["drhodt_as%d" % l] for l in lv), N,
                   f32("dt"), int(V["iter"]), steps, d)
        elif key == ("basic/time_scheme/adam_bashforth.cl", "postcorrector"):
            lv = range(1, 5)
            O.call("ab_postcorrector", O.ptrs(V["dudt_as%d" % l] for l in lv),
                   O.ptrs(V["drhodt_as%d" % l] for l in lv), V["dudt"], V["drhodt"],
                   O.ptrs(V["dudt_as%d_in" % l] for l in lv), O.ptrs(V["drhodt_as%d_in" % l] for l in lv), N, d)
        elif key == ("cfd/Boundary/BI/NoSlip.cl", "entry"):
            c("bi_noslip", D, self.ll(), V["iset"], V["imove"], V["r"], V["normal"], V["u"], V["rho"], V["m"],
              V["lap_u"], int(V["noslip_iset"]), f32("dr"))
        elif rel.startswith("cfd/ideal_gas/") and (rel, entry) in (
                ("cfd/ideal_gas/EOS.cl", "entry"), ("cfd/ideal_gas/Rates.cl", "entry"),
                ("cfd/ideal_gas/Sort.cl", "entry"), ("cfd/ideal_gas/TimeStep.cl", "entry"),
                ("cfd/ideal_gas/riemann/Rates.cl", "entry"), ("cfd/ideal_gas/symmetry/Mirror.cl", "set"),
                ("cfd/ideal_gas/riemann/Interactions.cl", "entry"),
                ("cfd/ideal_gas/time_scheme/euler.cl", "predictor"), ("cfd/ideal_gas/time_scheme/euler.cl", "corrector"),
                ("cfd/ideal_gas/time_scheme/improved_euler.cl", "predictor"),
                ("cfd/ideal_gas/time_scheme/improved_euler.cl", "corrector"),
                ("cfd/ideal_gas/time_scheme/midpoint.cl", "predictor"),
                ("cfd/ideal_gas/time_scheme/midpoint.cl", "midpoint"),
                ("cfd/ideal_gas/time_scheme/midpoint.cl", "relax"),
                ("cfd/ideal_gas/time_scheme/midpoint.cl", "corrector")):
            # the element-wise kernels of the ideal-gas presets (aqo_kernels.c, bit-identical to the scripts)
            if rel.endswith("EOS.cl"):
                c("ig_eos", V["iset"], V["imove"], V["rho"], V["eint"], V["p"], V["gamma"], N)
            elif rel == "cfd/ideal_gas/Rates.cl":
                c("ig_rates", V["imove"], V["rho"], V["p"], V["div_u"], V["deintdt"], N)
            elif rel.endswith("Sort.cl"):
                O.call("ig_sort", V["eint_in"], V["eint"], V["deintdt"], V["deintdt_in"], V["id_sorted"], N)
            elif rel.endswith("TimeStep.cl"):
                c("ig_timestep", D, V["dt_var"], V["imove"], V["iset"], V["u"], V["rho"], V["p"], N, f32("dt"),
                  f32("dt_min"), f32("courant"), V["div_u"], V["grad_p"], V["gamma"])
            elif rel.endswith("riemann/Interactions.cl"):
                c("ig_riemann_interactions", D, self.ll(), V["iset"], V["imove"], V["r"], V["u"], V["rho"], V["m"],
                  V["p"], V["grad_p"], V["div_u"], V["work_density"], V["gamma"])
            elif rel.endswith("symmetry/Mirror.cl"):
                c("ig_sym_set", V["mirror_src"], V["eint_in"], V["deintdt_in"], V["deintdt"], N)
            elif rel.endswith("riemann/Rates.cl"):
                c("ig_riemann_rates", V["imove"], V["work_density"], V["deintdt"], N)
            elif rel.endswith("time_scheme/euler.cl") and entry == "corrector":
                c("ig_euler_corrector", V["imove"], V["eint"], V["deintdt"], N, f32("dt"))
            elif rel.endswith("improved_euler.cl") and entry == "predictor":
                c("ig_ie_predictor", V["imove"], V["eint"], V["deintdt"], V["eint_in"], V["deintdt_in"], N, f32("dt"))
            elif rel.endswith("improved_euler.cl"):
                c("ig_ie_corrector", V["imove"], V["deintdt"], V["deintdt_in"], V["eint"], N, f32("dt"))
            elif entry == "predictor":
                c("ig_mp_predictor", V["eint"], V["deintdt"], V["eint_in"], V["deintdt_in"], N)
            elif entry == "midpoint":
                c("ig_mp_midpoint", V["imove"], V["eint_in"], V["deintdt"], V["eint"], N, f32("dt"))
            elif entry == "relax":
                c("ig_mp_relax", V["imove"], V["deintdt_in"], V["deintdt"], N, f32("relax_midpoint"))
            else:
                c("ig_mp_corrector", V["imove"], V["eint_in"], V["deintdt"], V["eint"], N, f32("dt"))
        elif rel == "cfd/Boundary/Symmetry/Mirror.cl":
            # preset cfd/symmetry.xml: detect / feed / set / sort / drop (aqo_kernels.c, bit-identical to the script)
            if entry == "detect":
                c("sym_detect", D, V["imove"], V["r_in"], V["imirror"], N, V["symmetry_r"], V["symmetry_n"])
            elif entry == "feed":
                O.call("sym_feed", V["imove"], V["iset"].view(np.int32), V["imirror"], V["imirror_invperm"],
                       V["mirror_src"], V["normal"], V["tangent"], V["r_in"], N, int(V["nbuffer"]), V["symmetry_r"],
                       V["symmetry_n"], d)
            elif entry == "set":
                c("sym_set", V["mirror_src"], V["m"], V["u_in"], V["dudt_in"], V["dudt"], V["rho_in"], V["drhodt_in"],
                  V["drhodt"], N, V["symmetry_n"], d)
            elif entry == "sort":
                c("sym_sort", V["mirror_src_in"], V["mirror_src"], V["id_sorted"], N)
            elif entry == "drop":
                c("sym_drop", V["imove"], V["r"], N, V["symmetry_r"], V["symmetry_n"], V["domain_max"], d)
            else:
                raise NotImplementedError("oracle interpreter: kernel %s::%s" % (rel, entry))
        elif key == ("cfd/Boundary/Portal/Shepard.cl", "entry"):
            c("portal_shepard", D, self.ll(), V["imove"], V["imirrored"], V["r"], V["rho"], V["m"], V["shepard"])
        elif key == ("cfd/Boundary/Portal/Interactions.cl", "entry"):
            lap = str(self.defs.get("__LAP_FORMULATION__", "__LAP_MONAGHAN__")).strip()
            lap = str(self.defs.get(lap, lap)).strip()
            if lap not in ("1", "2", "__LAP_MONAGHAN__", "__LAP_MORRIS__"):
                raise NotImplementedError("oracle interpreter: __LAP_FORMULATION__=" + lap)
            c("portal_interactions", D, self.ll(), V["imove"], V["imirrored"], V["r"], V["u"], V["rho"], V["m"],
              V["p"], V["grad_p"], V["lap_u"], V["div_u"], 1 if lap in ("2", "__LAP_MORRIS__") else 0)
        elif rel in ("cfd/Boundary/Inlet/Inlet.cl", "cfd/Boundary/Outlet/Outlet.cl", "cfd/Boundary/Portal/Mirror.cl"):
            # presets cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml: the element-wise kernels of the open
            # boundaries (aqo_kernels.c, bit-identical to the scripts)
            vec = lambda n: np.ascontiguousarray(V[n], np.float32)  # noqa: E731
            if rel.endswith("Inlet.cl") and entry == "feed":
                c("inlet_feed", D, V["imove"], V["iset"], V["r"], V["u"], V["dudt"], V["rho"], V["drhodt"], V["m"],
                  V["p"], V["refd"], N, int(V["nbuffer"]), f32("cs"), f32("p0"), vec("g"), f32("dr"), vec("inlet_r"),
                  vec("inlet_ru"), vec("inlet_rv"), np.ascontiguousarray(V["inlet_N"], np.uint32), vec("inlet_n"),
                  f32("inlet_U"), vec("inlet_rFS"), f32("inlet_R"), int(V["inlet_starving"]))
            elif rel.endswith("Inlet.cl") and entry == "rates":
                c("inlet_rates", V["imove"], V["r"], V["u"], V["dudt"], V["drhodt"], N, vec("inlet_r"),
                  f32("inlet_U"), vec("inlet_n"), d)
            elif rel.endswith("Outlet.cl") and entry == "rates":
                c("outlet_rates", V["imove"], V["iset"], V["r"], V["u"], V["rho"], V["p"], V["dudt"], V["dudt_in"],
                  V["drhodt"], V["drhodt_in"], V["refd"], N, f32("cs"), f32("p0"), vec("g"), vec("outlet_r"),
                  vec("outlet_n"), f32("outlet_U"), vec("outlet_rFS"), d)
            elif rel.endswith("Outlet.cl") and entry == "feed":
                c("outlet_feed", D, V["imove"], V["r_in"], N, vec("domain_max"), vec("outlet_r"), vec("outlet_n"))
            elif entry == "mirror":
                c("portal_mirror", D, V["r"], V["imirrored"], V["icell"], N, vec("portal_in_r"), vec("portal_out_r"),
                  vec("portal_n"), vec("r_min"), np.ascontiguousarray(V["n_cells"], np.uint32))
            elif entry == "unmirror":
                c("portal_unmirror", V["r"], V["imirrored"], N, vec("portal_in_r"), vec("portal_out_r"), d)
            elif entry == "teleport":
                c("portal_teleport", V["r"], N, vec("portal_in_r"), vec("portal_out_r"), vec("portal_n"), d)
            else:
                raise NotImplementedError("oracle interpreter: kernel %s::%s" % (rel, entry))
        elif rel == "aqua/MPIdeltaSPH.cl":
            # remote (halo) terms of MLS / delta-SPH: ours, not reference scripts (aqo_kernels.c)
            rl = O.make_ll(V["mpi_icell"], V["mpi_ihoc"], V["n_cells"], N)
            if entry == "mls":
                c("mpi_mls", D, rl, V["icell"], V["imove"], V["r"], V["mpi_r"], V["mpi_rho"], V["mpi_m"],
                  V["mls"], int(V["mls_imove"]))
            elif entry == "full_lapp":
                c("mpi_dsph_full_lapp", D, rl, V["icell"], V["imove"], V["r"], V["p"], V["mpi_r"], V["mpi_rho"],
                  V["mpi_m"], V["mpi_p"], V["lap_p_corr"], V["lap_p"])
            elif entry == "lapp_corr":
                c("mpi_dsph_lapp_corr", D, rl, V["icell"], V["imove"], V["r"], V["lap_p_corr"], V["mpi_r"],
                  V["mpi_rho"], V["mpi_m"], V["mpi_lap_p_corr"], V["lap_p"])
            elif entry == "copy_g":
                V["mpi_lap_p_corr"][:N] = V["lap_p_corr"][:N]
            elif entry == "sort_g":
                V["mpi_lap_p_corr"][V["mpi_id_sorted"][:N]] = V["mpi_lap_p_corr_in"][:N]
            else:
                raise NotImplementedError("oracle interpreter: kernel %s::%s" % (rel, entry))
        elif os.path.basename(rel) in ("bc.cl", "BlastRim.cl") and entry in ("set_fixed", "unset_fixed"):
            # the case-local script of examples/2D/shock_point (src/templates/bc.cl:44-66; tests/scripts/user/
            # BlastRim.cl is this repository's own wording of it): the rim is frozen while the time scheme runs
            f = np.float32
            if entry == "set_fixed":
                r = V["r"]
                length = np.sqrt((r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]).astype(f)).astype(f)
                band = f(f(f(1.5) * f(D.SUPPORT)) * f(D.H))
                V["imove"][length > f(f(V["R"]) - band)] = 0
            else:
                V["imove"][...] = 1
        elif rel.endswith("h_sensor.cl"):
            # examples/3D/spheric_testcase2_dambreak/src/templates/h_sensor.cl:1-60
            r, dr = V["r"], np.float32(V["dr"])
            x = r[:, 0] - np.float32(V["h_sensorx"])
            keep = (V["imove"] > 0) & ~((np.abs(x) > np.float32(2) * dr) | (np.abs(r[:, 1]) > np.float32(2) * dr))
            V["h_sensorz"][...] = np.where(keep, r[:, 2] + np.float32(0.5) * dr, np.float32(0))
        else:
            self._ref_kernel(rel, entry)

    def _ref_kernel(self, rel, entry, n=None):
        """Kernels without a plain-C restatement (cfd/MPI.cl, cfd/MPI/planes.cl,
        basic/SetBuffer.cl) run through the reference's OWN script compiled behind
        the shim (oracle/ref.py), arguments bound by name like Kernel.cpp:497-556."""
        from . import ref
        if self._ref is None:
            if not ref.available() and not ref.build():
                raise NotImplementedError("oracle interpreter: kernel %s::%s needs oracle/_ref" % (rel, entry))
            self._ref = ref.Ref(self.dims, float(self.V["h"]))
        if (rel, entry) not in self._ref.index:
            raise NotImplementedError("oracle interpreter: kernel %s::%s" % (rel, entry))
        names, kinds = self._ref.index[(rel, entry)]
        if n is None:   # Kernel::computeGlobalWorkSize (Kernel.cpp:558-594)
            n = max([self.V[k].shape[0] for k, kd in zip(names, kinds) if kd == "ptr"] + [1])
        self._ref.run(rel, entry, n, self.V)

    def _bie_eb(self):
        V = self.V
        O.call("bie_elastic_bounce", self.ll(), V["imove"], V["r_in"], V["normal"], V["m"],
               V["u_in"], V["dudt"], float(V["dt"]), float(self.dr_factor), self.dims)

    def _bie_pst(self):
        V = self.V
        O.call("bie_pst", self.ll(), V["imove"], V["r"], V["normal"], V["m"], V["rho"],
               float(self.defs["DIMS"]), float(self.dr_factor), self.dims)

    # -- I/O helpers
    def unsorted(self, name):
        return O.scatter(self.V[name], self.V["id"])
