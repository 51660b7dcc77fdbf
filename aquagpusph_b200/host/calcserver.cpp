// calcserver.cpp -- see calcserver.hpp
#include "calcserver.hpp"
#include "devloop.hpp"

#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>

namespace Aqua {
namespace CalcServer {

using InputOutput::ProblemSetup;
using InputOutput::Variable;
using InputOutput::Variables;

// ---------------------------------------------------------------- Tool base --
void Tool::check(int rc) const
{
    if (rc)
        throw std::runtime_error("Failure executing the tool \"" + _name +
                                 "\": " + aqc_last_error(_C->ctx()));
}

void Tool::execute()
{
    if (_once && _n_iters > 0)
        return;
    // AQUA_PROFILE_SYNC=1: per-tool device + host time (the stream is drained around every
    // tool, like the reference's per-tool Profile samples; Tool.cpp:296-310) -- diagnostics only
    static const bool prof_sync = getenv("AQUA_PROFILE_SYNC") != nullptr;
    if (prof_sync)
        aqc_sync(_C->ctx());
    const auto t0 = std::chrono::steady_clock::now();
    _execute();
    if (prof_sync)
        aqc_sync(_C->ctx());
    _elapsed_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    _n_iters++;
}

Variable* Tool::variable(const std::string& name, bool must_be_array, bool must_be_scalar) const
{
    Variable* v = _C->variables()->get(name);
    if (!v)
        throw std::runtime_error("The tool \"" + _name + "\" is asking the undeclared variable \"" +
                                 name + "\"");
    if (must_be_array && !v->isArray())
        throw std::runtime_error("The tool \"" + _name + "\" needs \"" + name +
                                 "\" to be an array, but it is a scalar");
    if (must_be_scalar && v->isArray())
        throw std::runtime_error("The tool \"" + _name + "\" needs \"" + name +
                                 "\" to be a scalar, but it is an array");
    return v;
}

// ------------------------------------------------------------------- Kernel --
Kernel::Kernel(CalcServer* C, const std::string& name, const std::string& path,
               const std::string& entry, const std::string& n, bool once)
  : Tool(C, name, once), _path(path), _entry(entry), _n(n)
{
}

void Kernel::setup()
{
    _kid = aqc_kernel_lookup(_path.c_str(), _entry.c_str(), _C->dims());
    // Under __LAP_FORMULATION__ = __LAP_MORRIS__ only cfd/Interactions.cl has a hand-written build of that
    // branch: the other scripts whose Laplacian term depends on it run as the scripts themselves
    if (_kid >= 0 && _C->lapMorris()) {
        auto ends = [&](const char* tail) {
            const size_t n = strlen(tail);
            return _path.size() >= n && _path.compare(_path.size() - n, n, tail) == 0;
        };
        if (ends("cfd/Boundary/BI/LapU.cl") || ends("cfd/Boundary/BI/NoSlip.cl") ||
            (ends("cfd/MPI.cl") && _entry == "interactions"))
            _kid = -1;
    }
    if (_kid < 0) {
        // not a hand-written kernel: the script itself, compiled at run time like the reference compiles
        // every script (Kernel.cpp:354-420) -- here by NVRTC for sm_100a (csrc/clc.cu)
        std::vector<std::string> defs;
        for (auto& d : _C->definitions())
            defs.push_back("-D" + d.first + (d.second.empty() ? "" : "=" + d.second));
        std::vector<const char*> dp;
        for (auto& d : defs)
            dp.push_back(d.c_str());
        _kid = aqc_script_compile(_C->ctx(), _path.c_str(), _entry.c_str(), _C->dims(),
                                  _C->sim_data().settings.base_path.c_str(), dp.data(), (int)dp.size());
        if (_kid < 0)
            throw std::runtime_error("The tool \"" + name() + "\" asks for the script \"" + _path +
                                     "\" entry point \"" + _entry + "\", which is not in the CUDA kernel "
                                     "registry and could not be compiled at run time: " + aqc_last_error(_C->ctx()));
        log(L_INFO, "The tool \"" + name() + "\" runs the script \"" + _path + "\"::" + _entry +
                        " compiled at run time (NVRTC)\n");
    }
    const int na = aqc_kernel_nargs(_kid);
    const aqc_arg_info* info = aqc_kernel_args(_kid);
    Variables* vars = _C->variables();
    for (int k = 0; k < na; k++) {
        Variable* v = vars->get(info[k].name);
        if (!v)
            throw std::runtime_error("The tool \"" + name() + "\" requires the undeclared variable \"" +
                                     info[k].name + "\"");
        const bool want_array = info[k].kind != AQC_ARG_SCALAR;
        if (want_array != v->isArray())
            throw std::runtime_error("The tool \"" + name() + "\": variable \"" + v->name() +
                                     "\" is a " + (v->isArray() ? "array" : "scalar") +
                                     " but the kernel expects a " + (want_array ? "array" : "scalar"));
        // type check like Kernel.cpp:518-540 (size_t/svec aliases resolved)
        std::string kt = info[k].type;
        if (!vars->isSameType(kt, v->type()))
            throw std::runtime_error("The tool \"" + name() + "\": variable \"" + v->name() +
                                     "\" has type \"" + v->type() + "\" but the kernel expects \"" +
                                     kt + "\"");
        _vars.push_back(v);
        _kinds.push_back(info[k].kind);
    }
}

bool Kernel::dependencies(std::vector<Variable*>& in, std::vector<Variable*>& out) const
{
    // Kernel.cpp:464-556: const / non-global arguments are inputs, the rest outputs
    for (size_t k = 0; k < _vars.size(); k++)
        (_kinds[k] == AQC_ARG_ARRAY_OUT ? out : in).push_back(_vars[k]);
    return true;
}

void Kernel::_execute()
{
    if (_leader)
        return; // computed by the fused launch of the group's leader
    if (_fused_id >= 0) {
        std::vector<void*> args;
        for (auto k : _group)
            for (auto v : k->_vars)
                args.push_back(v->isArray() ? v->dptr() : v->get());
        check(aqc_launch_fused(_C->ctx(), _fused_id, args.data(), (int)args.size()));
        return;
    }
    const size_t N = globalSize();
    std::vector<void*> args(_vars.size());
    for (size_t k = 0; k < _vars.size(); k++)
        args[k] = _vars[k]->isArray() ? _vars[k]->dptr() : _vars[k]->get();
    check(aqc_launch(_C->ctx(), _kid, N, args.data(), (int)args.size()));
}

// --------------------------------------------------------------------- Copy --
void Copy::setup()
{
    _in = variable(_in_name, true);
    _out = variable(_out_name, true);
    if (!_C->variables()->isSameType(_in->type(), _out->type()))
        throw std::runtime_error("The tool \"" + name() + "\": mismatching types \"" + _in->type() +
                                 "\" and \"" + _out->type() + "\"");
    if (_in->size() != _out->size())
        throw std::runtime_error("The tool \"" + name() + "\": mismatching lengths");
}

void Copy::_execute() { check(aqc_memcpy_d2d(_C->ctx(), _out->dptr(), _in->dptr(), _in->size())); }

// ---------------------------------------------------------------------- Set --
// OpenCL literals the presets use as values / identities (Set.cl.in, Reduction.hcl.in,
// types/{2D,3D}.h): returns false when `s` is not one of them.
static bool literalValue(const Variables* vars, const std::string& type, std::string s,
                         std::vector<char>& out)
{
    const size_t ts = vars->typeToBytes(type);
    const unsigned n = ts / 4;
    std::vector<float> f(n, 0.f);
    s = trimCopy(s);
    auto splat3 = [&](float v, bool all) {
        for (unsigned c = 0; c < n; c++)
            f[c] = v;
        if (!all && vars->dims() == 3 && n == 4)
            f[3] = 0.f;
    };
    bool neg = false;
    if (startswith(s, "-")) {
        neg = true;
        s = trimCopy(s.substr(1));
    }
    if (s == "VEC_ZERO" || s == "MAT_ZERO") splat3(0.f, true);
    else if (s == "VEC_ONE") splat3(1.f, false);
    else if (s == "VEC_ALL_ONE" || s == "MAT_ALL_ONE") splat3(1.f, true);
    else if (s == "VEC_INFINITY") splat3(INFINITY, false);
    else if (s == "VEC_ALL_INFINITY") splat3(INFINITY, true);
    else if (s == "VEC_NEG_INFINITY") splat3(-INFINITY, false);
    else if (s == "VEC_ALL_NEG_INFINITY") splat3(-INFINITY, true);
    else if (s == "INFINITY") splat3(INFINITY, true);
    else if (s == "MAT_EYE" || s == "MAT_ALL_EYE") {
        const unsigned d = (n == 16) ? 4 : 2;
        for (unsigned a = 0; a < d; a++)
            f[a * d + a] = (n == 16 && a == 3 && s == "MAT_EYE") ? 0.f : 1.f;
    } else
        return false;
    if (neg)
        for (auto& v : f)
            v = -v;
    out.resize(ts);
    memcpy(out.data(), f.data(), ts);
    return true;
}

// Evaluate `value` as an element of array type `type`: an expression (a single
// one is widened to every component, like OpenCL scalar->vector conversion), a
// cast literal "((vec4)(0.f))", or one of the macros above.
static bool elementValue(Variables* vars, const std::string& type, const std::string& value,
                         std::vector<char>& out)
{
    std::string t = trimCopy(type);
    if (!t.empty() && t.back() == '*')
        t.pop_back();
    const size_t ts = vars->typeToBytes(t);
    std::string v = trimCopy(value);
    // strip "((type)(...))" casts
    for (int guard = 0; guard < 4; guard++) {
        if (v.size() > 2 && v.front() == '(' && v.back() == ')') {
            int depth = 0;
            bool wraps = true;
            for (size_t i = 0; i < v.size(); i++) {
                if (v[i] == '(') depth++;
                if (v[i] == ')') depth--;
                if (depth == 0 && i + 1 < v.size()) { wraps = false; break; }
            }
            if (wraps) { v = trimCopy(v.substr(1, v.size() - 2)); continue; }
        }
        if (v.size() > 2 && v.front() == '(') {
            const size_t c = v.find(')');
            const std::string inner = trimCopy(v.substr(1, c - 1));
            if (c != std::string::npos && vars->typeToBytes(inner) && c + 1 < v.size()) {
                v = trimCopy(v.substr(c + 1));
                continue;
            }
        }
        break;
    }
    if (literalValue(vars, t, v, out))
        return true;
    out.assign(ts, 0);
    const unsigned n = vars->typeToN(t) == 3 ? 4 : vars->typeToN(t);
    try {
        auto parts = split_formulae(v);
        if (parts.size() == 1 && n > 1) {
            // widen a scalar expression
            const size_t cs = ts / n;
            std::string base = t;
            std::vector<char> one(8, 0);
            // component type: float / int / unsigned int
            std::string ct = "float";
            if (startswith(vars->typeAlias(t), "uivec")) ct = "unsigned int";
            else if (startswith(vars->typeAlias(t), "ivec")) ct = "int";
            vars->solve(ct, v, one.data());
            for (unsigned c = 0; c < n; c++)
                memcpy(out.data() + c * cs, one.data(), cs);
        } else {
            vars->solve(t, v, out.data());
        }
    } catch (std::exception&) {
        return false;
    }
    return true;
}

void Set::setup()
{
    _var = variable(_var_name, true);
    // a value that only uses constants can be evaluated once; otherwise per step
    std::vector<char> tmp;
    if (!elementValue(_C->variables(), _var->type(), _value, tmp))
        throw std::runtime_error("The tool \"" + name() + "\": cannot evaluate the value \"" +
                                 _value + "\" for the array \"" + _var_name + "\"");
}

void Set::_execute()
{
    if (!elementValue(_C->variables(), _var->type(), _value, _data))
        throw std::runtime_error("The tool \"" + name() + "\": cannot evaluate \"" + _value + "\"");
    const size_t ts = _var->typesize();
    check(aqc_fill(_C->ctx(), _var->dptr(), _var->length(), ts, _data.data()));
}

// --------------------------------------------------------- ScalarExpression --
void ScalarExpression::setup()
{
    _value.assign(_C->variables()->typeToBytes(_type), 0);
    // dependencies must exist (SetScalar.cpp:205-233)
    _C->variables()->exprVariables(_expr);
}

void ScalarExpression::solve()
{
    try {
        _C->variables()->solve(_type, _expr, _value.data());
    } catch (std::exception& e) {
        throw std::runtime_error("The tool \"" + name() + "\" failed evaluating \"" + _expr +
                                 "\": " + e.what());
    }
}

void SetScalar::setup()
{
    _var = variable(_var_name, false, true);
    _type = _var->type();
    ScalarExpression::setup();
}

void SetScalar::_execute()
{
    solve();
    _var->set(_value.data());
    _C->variables()->populate(_var);
}

void Assert::_execute()
{
    solve();
    int r;
    memcpy(&r, _value.data(), sizeof(int));
    if (!r)
        throw std::runtime_error("Assertion error. The expression \"" + _expr + "\" of the tool \"" +
                                 name() + "\" is false");
}

// -------------------------------------------------------------- Conditional --
void Conditional::setup()
{
    auto& tools = _C->tools();
    int i = id_in_pipeline();
    if (i < 0)
        throw std::runtime_error("Invalid tool \"" + name() + "\"");
    int scope = 1;
    while ((size_t)i < tools.size() - 1) {
        i++;
        scope += tools[i]->scope_modifier();
        if (!scope)
            break;
    }
    if (scope > 0)
        throw std::runtime_error("Unbalanced scope opened by the tool \"" + name() + "\"");
    _ending_tool = ((size_t)i == tools.size() - 1) ? nullptr : tools[i + 1].get();
    ScalarExpression::setup();
}

void Conditional::_execute()
{
    solve();
    int r;
    memcpy(&r, _value.data(), sizeof(int));
    _result = r != 0;
}

Tool* Conditional::next_tool() { return _result ? Tool::next_tool() : _ending_tool; }

Tool* If::next_tool()
{
    Tool* next = Conditional::next_tool();
    // Conditional.cpp:120-136: the matching End hands control back to this tool; the
    // flag is flipped so that the second visit falls through to the ending tool
    _result = !_result;
    return next;
}

void If::_execute()
{
    if (_result)
        Conditional::_execute();
}

void End::setup()
{
    auto& tools = _C->tools();
    int i = id_in_pipeline();
    if (i < 0)
        throw std::runtime_error("Invalid tool \"" + name() + "\"");
    int scope = 1;
    while (i > 0) {
        i--;
        scope -= tools[i]->scope_modifier();
        if (!scope)
            break;
    }
    if (scope > 0)
        throw std::runtime_error("The tool \"" + name() + "\" closes a scope that was never opened");
    next_tool(tools[i].get());
}

// ---------------------------------------------------------------- Reduction --
void Reduction::setup()
{
    _in = variable(_in_name, true);
    _out = variable(_out_name, false, true);
    Variables* vars = _C->variables();
    if (!vars->isSameType(_in->type(), _out->type()))
        throw std::runtime_error("The tool \"" + name() + "\": mismatching input and output types \"" +
                                 _in->type() + "\" / \"" + _out->type() + "\"");
    // normalise the OpenCL snippet "c = f(a, b);"
    std::string op;
    for (char c : _operation)
        if (!isspace((unsigned char)c) && c != ';')
            op.push_back(c);
    if (op == "c=a+b" || op == "c=b+a") _op = AQC_OP_SUM;
    else if (op == "c=min(a,b)" || op == "c=min(b,a)" || op == "c=(a<b)?a:b" || op == "c=(a>b)?b:a" ||
             op == "c=(a<=b)?a:b" || op == "c=(a>=b)?b:a" || op == "c=fmin(a,b)") _op = AQC_OP_MIN;
    else if (op == "c=max(a,b)" || op == "c=max(b,a)" || op == "c=(a<b)?b:a" || op == "c=(a>b)?a:b" ||
             op == "c=(a<=b)?b:a" || op == "c=(a>=b)?a:b" || op == "c=fmax(a,b)") _op = AQC_OP_MAX;
    else
        throw std::runtime_error("The tool \"" + name() + "\": unsupported reduction operation \"" +
                                 trimCopy(_operation) + "\" (supported: a + b, min, max and their "
                                 "ternary forms)");
    const char k = _in->kind();
    const unsigned n = _in->ncomp();
    if (k == 'f' && n == 1) _atype = AQC_T_F32;
    else if (k == 'f' && n == 2) _atype = AQC_T_VEC2;
    else if (k == 'f' && n == 4) _atype = AQC_T_VEC4;
    else if (k == 'u' && n == 1) _atype = AQC_T_U32;
    else if (k == 'i' && n == 1) _atype = AQC_T_I32;
    else
        throw std::runtime_error("The tool \"" + name() + "\": unsupported reduction type \"" +
                                 _in->type() + "\"");
    if (!elementValue(vars, _in->type(), _null, _identity))
        throw std::runtime_error("The tool \"" + name() + "\": cannot evaluate the null value \"" +
                                 _null + "\"");
}

void Reduction::_execute()
{
    char res[16] = { 0 };
    check(aqc_reduce(_C->ctx(), _op, _atype, _in->dptr(), _in->length(), nullptr, res));
    // fold the user's identity in (e.g. VEC_INFINITY has w = 0, Reduction.hcl.in:90)
    const unsigned n = _in->ncomp();
    for (unsigned c = 0; c < n; c++) {
        if (_in->kind() == 'f') {
            float a, b;
            memcpy(&a, res + 4 * c, 4);
            memcpy(&b, _identity.data() + 4 * c, 4);
            const float r = _op == AQC_OP_SUM ? a + b : (_op == AQC_OP_MIN ? fminf(a, b) : fmaxf(a, b));
            memcpy(res + 4 * c, &r, 4);
        } else if (_in->kind() == 'u') {
            uint32_t a, b;
            memcpy(&a, res + 4 * c, 4);
            memcpy(&b, _identity.data() + 4 * c, 4);
            const uint32_t r = _op == AQC_OP_SUM ? a + b : (_op == AQC_OP_MIN ? std::min(a, b) : std::max(a, b));
            memcpy(res + 4 * c, &r, 4);
        } else {
            int32_t a, b;
            memcpy(&a, res + 4 * c, 4);
            memcpy(&b, _identity.data() + 4 * c, 4);
            const int32_t r = _op == AQC_OP_SUM ? a + b : (_op == AQC_OP_MIN ? std::min(a, b) : std::max(a, b));
            memcpy(res + 4 * c, &r, 4);
        }
    }
    _out->set(res);
    _C->variables()->populate(_out);
}

// ----------------------------------------------------------------- LinkList --
LinkList::LinkList(CalcServer* C, const std::string& name, const ProblemSetup::Tool& t, bool once)
  : Tool(C, name, once)
  , _in_name(t.get("in")), _min_name(t.get("min")), _max_name(t.get("max"))
  , _ihoc_name(t.get("ihoc")), _icell_name(t.get("icell")), _ncells_name(t.get("n_cells"))
  , _perm_name(t.get("perm")), _inv_name(t.get("inv_perm"))
  , _recompute(toLowerCopy(t.get("recompute_grid")) != "false")
  , _depends_txt(t.get("depends"))
{
    const std::string sorter = t.get("sorter");
    if (!sorter.empty() && sorter != "radix-sort" && sorter != "bitonic")
        throw std::runtime_error("The tool \"" + name + "\": unknown sorter \"" + sorter + "\"");
    // sorter="bitonic" is served by the same stable device sort
}

void LinkList::setup()
{
    _in = variable(_in_name, true);
    _min = variable(_min_name, false, true);
    _max = variable(_max_name, false, true);
    _ihoc = variable(_ihoc_name, true);
    _icell = variable(_icell_name, true);
    _ncells = variable(_ncells_name, false, true);
    _perm = variable(_perm_name, true);
    _inv = variable(_inv_name, true);
    for (auto& d : split(_depends_txt))
        if (!trimCopy(d).empty())
            _depends.push_back(variable(trimCopy(d), true));
    _N = variable("N", false, true);
    _support = variable("support", false, true);
    _h = variable("h", false, true);
    _ihoc->reallocatable(true); // LinkList.cpp:179-182
    const float cell = *(float*)_support->get() * *(float*)_h->get();
    if (!cell)
        throw std::runtime_error("Zero cell length detected in the tool \"" + name() + "\"");
}

void LinkList::_execute()
{
    static const bool verify = getenv("AQC_MPI_VERIFY") && atoi(getenv("AQC_MPI_VERIFY"));
    std::vector<uint32_t> old_icell, old_perm;
    if (!_depends.empty()) {
        if (_watch < 0)
            _watch = aqc_watch_create(_C->ctx());
        if (!aqc_watch_dirty(_C->ctx(), _watch) && _built_in == _in->dptr() && _built_n == _in->length()) {
            if (!verify) {
                _skipped++;
                return;
            }
            // tests: build anyway and require the outputs to be what they were
            old_icell.resize(_icell->length());
            old_perm.resize(_perm->length());
            check(aqc_memcpy_d2h(_C->ctx(), old_icell.data(), _icell->dptr(), _icell->size(), 1));
            check(aqc_memcpy_d2h(_C->ctx(), old_perm.data(), _perm->dptr(), _perm->size(), 1));
        }
    }
    float rmin[4] = { 0, 0, 0, 0 }, rmax[4] = { 0, 0, 0, 0 };
    memcpy(rmin, _min->get(), _min->typesize());
    memcpy(rmax, _max->get(), _max->typesize());
    aqc_usize nc[4];
    aqc_usize* ihoc = (aqc_usize*)_ihoc->dptr();
    size_t cap = _ihoc->length();
    check(aqc_linklist_build(_C->ctx(), _in->dptr(), (aqc_usize)_in->length(), _C->dims(),
                             *(float*)_support->get(), *(float*)_h->get(), _recompute ? 1 : 0,
                             rmin, rmax, nc, (aqc_usize*)_icell->dptr(), &ihoc, &cap,
                             (aqc_usize*)_perm->dptr(), (aqc_usize*)_inv->dptr()));
    if (ihoc != _ihoc->dptr() || cap != _ihoc->length())
        _ihoc->reset(ihoc, cap);
    if (_recompute) {
        _min->set(rmin);
        _max->set(rmax);
        _C->variables()->populate(_min);
        _C->variables()->populate(_max);
    }
    _ncells->set(nc);
    _C->variables()->populate(_ncells);
    if (!old_icell.empty()) {
        std::vector<uint32_t> now_icell(old_icell.size()), now_perm(old_perm.size());
        check(aqc_memcpy_d2h(_C->ctx(), now_icell.data(), _icell->dptr(), _icell->size(), 1));
        check(aqc_memcpy_d2h(_C->ctx(), now_perm.data(), _perm->dptr(), _perm->size(), 1));
        if (now_icell != old_icell || now_perm != old_perm)
            throw std::runtime_error("The tool \"" + name() + "\" would have skipped a build whose result "
                                     "changed: its depends list is incomplete");
        _skipped++;
    }
    if (!_depends.empty()) {
        std::vector<const void*> ptrs;
        std::vector<size_t> bytes;
        for (auto v : _depends) {
            ptrs.push_back(v->dptr());
            bytes.push_back(v->length() * v->typesize());
        }
        check(aqc_watch_reset(_C->ctx(), _watch, (int)ptrs.size(), ptrs.data(), bytes.data()));
        _built_in = _in->dptr();
        _built_n = _in->length();
    }
}

// ---------------------------------------------------------------- RadixSort --
void RadixSort::setup()
{
    _var = variable(_var_name, true);
    _perm = variable(_perm_name, true);
    _inv = variable(_inv_name, true);
    for (auto v : { _var, _perm, _inv })
        if (v->kind() != 'u' || v->ncomp() != 1)
            throw std::runtime_error("The tool \"" + name() + "\": \"" + v->name() +
                                     "\" must be an unsigned int / size_t array");
}

void RadixSort::_execute()
{
    aqc_usize key_max = 0; // RadixSort.cpp:152-176: only "icell" has a tight bound
    if (_var_name == "icell") {
        aqc_usize nc[4];
        memcpy(nc, variable("n_cells", false, true)->get(), sizeof(nc));
        key_max = nc[3];
    }
    check(aqc_radix_sort(_C->ctx(), (aqc_usize*)_var->dptr(), (aqc_usize)_var->length(), key_max,
                         (aqc_usize*)_perm->dptr(), (aqc_usize*)_inv->dptr()));
}

// ------------------------------------------------------------------- UnSort --
void UnSort::setup()
{
    _in = variable(_in_name, true);
    _out = variable(_out_name, true);
    _perm = variable(_perm_name, true);
}

void UnSort::_execute()
{
    const void* src[1] = { _in->dptr() };
    void* dst[1] = { _out->dptr() };
    const size_t eb[1] = { _in->typesize() };
    check(aqc_scatter_fields(_C->ctx(), (const aqc_usize*)_perm->dptr(), (aqc_usize)_in->length(), 1,
                             src, dst, eb));
}

// ------------------------------------------------------------------ MPISync --
void MPISync::setup()
{
    // MPISync::variables (MPISync.cpp:232-298): mask must be size_t*, fields arrays of its length
    _mask = variable(_mask_name, true);
    if (!_C->variables()->isSameType(_mask->type(), "size_t*"))
        throw std::runtime_error("The tool \"" + name() + "\" is asking the variable \"" + _mask_name +
                                 "\", which has an invalid type (\"size_t*\" was expected)");
    for (auto& f : split(_fields_txt)) {
        if (trimCopy(f).empty())
            continue;
        InputOutput::Variable* v = variable(trimCopy(f), true);
        if (v->length() != _mask->length())
            throw std::runtime_error("Wrong variable length in the tool \"" + name() + "\": \"" +
                                     _mask_name + "\" has length " + std::to_string(_mask->length()) +
                                     ", \"" + v->name() + "\" has length " + std::to_string(v->length()));
        _fields.push_back(v);
    }
    // processes="" -> every other rank (CalcServer.cpp:354-374)
    for (auto& p : split(_procs_txt)) {
        if (trimCopy(p).empty())
            continue;
        unsigned v = 0;
        _C->variables()->solve("unsigned int", p, &v);
        if ((int)v >= _C->mpi_size())
            throw std::runtime_error("The tool \"" + name() + "\": process " + std::to_string(v) +
                                     " does not exist");
        _procs.push_back(v);
    }
    for (auto& d : split(_depends_txt))
        if (!trimCopy(d).empty())
            _depends.push_back(variable(trimCopy(d), true));
}

void MPISync::_execute()
{
    if (_C->mpi_size() <= 1)
        return; // MPISync.cpp:186-187
    std::vector<void*> ptrs;
    std::vector<size_t> eb;
    for (auto v : _fields) {
        ptrs.push_back(v->dptr());
        eb.push_back(v->typesize());
    }
    if (_plan < 0 && !_depends.empty())
        _plan = aqc_mpi_sync_plan(_C->ctx());
    std::vector<const void*> dptrs;
    std::vector<size_t> dbytes;
    for (auto v : _depends) {
        dptrs.push_back(v->dptr());
        dbytes.push_back(v->length() * v->typesize());
    }
    check(aqc_mpi_sync_ex(_C->ctx(), _plan, (aqc_usize*)_mask->dptr(), (aqc_usize)_mask->length(),
                          (int)ptrs.size(), ptrs.data(), eb.data(), (int)_procs.size(),
                          _procs.empty() ? nullptr : _procs.data(), nullptr, (int)dptrs.size(),
                          dptrs.data(), dbytes.data()));
}

// --------------------------------------------------------------- PythonTool --
namespace {
ScriptRunnerFn g_script_runner = nullptr;
void* g_script_user = nullptr;
}
void setScriptRunner(ScriptRunnerFn fn, void* user)
{
    g_script_runner = fn;
    g_script_user = user;
}

void PythonTool::setup()
{
    if (_path.empty())
        throw std::runtime_error("The tool \"" + name() + "\" names no script");
    if (!g_script_runner)
        throw std::runtime_error("The tool \"" + name() + "\" (type python, script \"" + _path +
                                 "\") needs a script runner: this host does not embed an interpreter, "
                                 "the driving process registers one with aqh_set_script_runner");
}

void PythonTool::_execute()
{
    // the script reads variables through aqh_scalar_get / aqh_array_download: everything enqueued so
    // far must have landed (the reference's getters wait for the variable's event the same way)
    check(aqc_sync(_C->ctx()));
    if (g_script_runner(g_script_user, name().c_str(), _path.c_str()))
        throw std::runtime_error("Python execution error in the tool \"" + name() + "\"");
}

// ------------------------------------------------------------- MPIAllReduce --
void MPIAllReduce::setup()
{
    _var = variable(_var_name, false, true);
    const std::string op = toLowerCopy(trimCopy(_op_txt));
    if (op == "min") _op = AQC_OP_MIN;
    else if (op == "max") _op = AQC_OP_MAX;
    else if (op == "sum" || op == "add") _op = AQC_OP_SUM;
    else
        throw std::runtime_error("The tool \"" + name() + "\": unknown operation \"" + _op_txt + "\"");
    if (_var->compsize() != 4 || _var->typesize() > 64)
        throw std::runtime_error("The tool \"" + name() + "\": \"" + _var_name +
                                 "\" must be made of 32-bit components");
    _type = _var->kind() == 'f' ? AQC_T_F32 : (_var->kind() == 'u' ? AQC_T_U32 : AQC_T_I32);
    _count = _var->ncomp();
}

void MPIAllReduce::_execute()
{
    if (_C->mpi_size() <= 1)
        return;
    std::vector<char> v(_var->typesize());
    memcpy(v.data(), _var->get(), v.size());
    check(aqc_allreduce_host(_C->ctx(), _op, _type, v.data(), _count));
    _var->set(v.data());
    _C->variables()->populate(_var);
}

// ------------------------------------------------------------------- Report --
Report::Report(CalcServer* C, const std::string& name, const std::string& kind,
               const ProblemSetup::Tool& t, bool once)
  : Tool(C, name, once), _kind(kind), _fields(t.get("fields")), _path(t.get("path"))
{
}

Report::~Report()
{
    if (_f)
        fclose(_f);
}

void Report::setup()
{
    if (_kind == "performance" || _kind == "particles")
        return;
    for (auto f : split(replaceAllCopy(_fields, " ", ","))) {
        if (f.empty())
            continue;
        Variable* v = _C->variables()->get(f);
        if (!v)
            throw std::runtime_error("The report \"" + name() + "\" is asking the undeclared variable \"" +
                                     f + "\"");
        _vars.push_back(v);
    }
    if (_kind == "file") {
        const std::string p = formatPath(_path, _C->mpi_rank());
        _f = fopen(p.c_str(), "w");
        if (!_f)
            throw std::runtime_error("The report \"" + name() + "\" cannot write \"" + p + "\"");
        fprintf(_f, "#");
        for (auto v : _vars)
            fprintf(_f, " %s", v->name().c_str());
        fprintf(_f, "\n");
    }
}

void Report::_execute()
{
    if (_kind == "screen") {
        if (logLevel() > L_INFO)
            return;
        std::string s = name() + ":";
        for (auto v : _vars)
            s += " " + v->name() + "=" + v->asString();
        log(L_INFO, s);
    } else if (_kind == "file") {
        for (auto v : _vars)
            fprintf(_f, "%s ", v->asString().c_str());
        fprintf(_f, "\n");
    } else if (_kind == "dump") {
        // Reports/Dump.cpp:90-175: one line per element, fields side by side
        const std::string p = formatPath(_path, _C->mpi_rank());
        FILE* f = fopen(p.c_str(), "w");
        if (!f)
            throw std::runtime_error("The report \"" + name() + "\" cannot write \"" + p + "\"");
        std::vector<std::vector<char>> host;
        size_t n = 0;
        for (auto v : _vars) {
            host.emplace_back(v->size());
            _C->download(v->name(), host.back().data());
            n = std::max(n, v->length());
        }
        for (size_t i = 0; i < n; i++) {
            for (size_t k = 0; k < _vars.size(); k++) {
                Variable* v = _vars[k];
                if (i >= v->length())
                    continue;
                const char* e = host[k].data() + i * v->typesize();
                for (unsigned c = 0; c < v->ncomp(); c++) {
                    if (v->kind() == 'f') fprintf(f, "%.9g ", *(const float*)(e + 4 * c));
                    else if (v->kind() == 'u') fprintf(f, "%u ", *(const uint32_t*)(e + 4 * c));
                    else fprintf(f, "%d ", *(const int32_t*)(e + 4 * c));
                }
            }
            fprintf(f, "\n");
        }
        fclose(f);
    }
}

// -------------------------------------------------------------- TimeManager --
TimeManager::TimeManager(CalcServer* C, const ProblemSetup& sd)
{
    Variables* vars = C->variables();
    const std::map<std::string, std::string> types{ { "t", "float" }, { "dt", "float" },
        { "iter", "unsigned int" }, { "frame", "unsigned int" }, { "end_t", "float" },
        { "end_iter", "unsigned int" }, { "end_frame", "unsigned int" } };
    for (auto& kv : types)
        if (!vars->get(kv.first) || vars->get(kv.first)->type() != kv.second)
            throw std::runtime_error("Expected a variable \"" + kv.first + "\" of type \"" + kv.second + "\"");
    _time = (float*)vars->get("t")->get();
    _dt = (float*)vars->get("dt")->get();
    _step = (unsigned*)vars->get("iter")->get();
    _frame = (unsigned*)vars->get("frame")->get();
    _time_max = (float*)vars->get("end_t")->get();
    _steps_max = (unsigned*)vars->get("end_iter")->get();
    _frames_max = (unsigned*)vars->get("end_frame")->get();
    typedef ProblemSetup::TimeOpts T;
    if (sd.time_opts.sim_end_mode & T::FRAME_MODE) *_frames_max = sd.time_opts.sim_end_frame;
    if (sd.time_opts.sim_end_mode & T::ITER_MODE) *_steps_max = sd.time_opts.sim_end_step;
    if (sd.time_opts.sim_end_mode & T::TIME_MODE) *_time_max = sd.time_opts.sim_end_time;
    if (sd.time_opts.output_mode & T::IPF_MODE) _output_ipf = (int)sd.time_opts.output_ipf;
    if (sd.time_opts.output_mode & T::FPS_MODE) _output_fps = sd.time_opts.output_fps;
    for (auto n : { "end_t", "end_iter", "end_frame" })
        vars->populate(n);
    _output_time = *_time;
    _output_step = *_step;
}

bool TimeManager::mustStop()
{
    return (*_time >= *_time_max) || (*_step >= *_steps_max) || (*_frame >= *_frames_max);
}

bool TimeManager::mustPrintOutput()
{
    if (*_time < 0.f) {
        *_step = 0;
        return false;
    }
    if (((_output_fps >= 0.f) || (_output_ipf >= 0)) && (*_frame == 0) && (*_step == 1)) {
        _output_time = *_time;
        _output_step = *_step;
        *_frame += 1;
        return true;
    }
    if ((_output_fps > 0.f) && (*_time - _output_time >= 1.f / _output_fps)) {
        _output_time += 1.f / _output_fps;
        _output_step = *_step;
        *_frame += 1;
        return true;
    }
    if ((_output_ipf > 0) && ((int)(*_step - _output_step) >= _output_ipf)) {
        _output_time = *_time;
        _output_step = *_step;
        *_frame += 1;
        return true;
    }
    return mustStop();
}

// --------------------------------------------------------------- CalcServer --
CalcServer::CalcServer(ProblemSetup& sd, int device, int mpi_rank, int mpi_size)
  : _sim_data(sd), _mpi_rank(mpi_rank), _mpi_size(mpi_size)
{
    // device selection: <Device> entries are indexed by rank (CalcServer.cpp:1163-1176);
    // "platform" has no meaning for CUDA, "device" is the CUDA ordinal
    if (device < 0) {
        device = 0;
        if (!sd.settings.devices.empty()) {
            const auto& d = sd.settings.devices[std::min<size_t>(mpi_rank, sd.settings.devices.size() - 1)];
            device = (int)d.device;
            if (d.addr_bits != 32)
                throw std::runtime_error("addr_bits=\"64\" devices are not supported: indices are 32 bit");
        }
    }
    if (aqc_ctx_create(device, &_ctx))
        throw std::runtime_error(std::string("Cannot create the CUDA context: ") + aqc_last_error(nullptr));
    {
        // every array is written through libaquacuda here, so the pair-mask cache of the neighbour
        // sweeps is safe (aquacuda.h); AQC_PAIR_CACHE=0 turns it off for A/B measurements
        const char* e = getenv("AQC_PAIR_CACHE");
        aqc_pairs_cache_enable(_ctx, (e && atoi(e) == 0) ? 0 : 1);
    }
    _vars = std::make_unique<Variables>(sd.dims, _ctx);

    size_t N = 0;
    for (auto& s : sd.sets)
        N += s->n;
    const size_t n_radix = nextPowerOf2(roundUp<size_t>(N, 64 * 16)); // _ITEMS * _GROUPS
    auto U = [](size_t v) { return std::to_string(v); };
    // default scalars (CalcServer.cpp:178-219)
    _vars->registerVariable("mpi_rank", "unsigned int", "", U(mpi_rank));
    _vars->registerVariable("mpi_size", "unsigned int", "", U(mpi_size));
    _vars->registerVariable("dims", "unsigned int", "", U(sd.dims));
    _vars->registerVariable("t", "float", "", "0");
    _vars->registerVariable("dt", "float", "", "0");
    _vars->registerVariable("iter", "unsigned int", "", "0");
    _vars->registerVariable("frame", "unsigned int", "", "0");
    {
        std::ostringstream v;
        v << std::numeric_limits<float>::max() / 2.f;
        _vars->registerVariable("end_t", "float", "", v.str());
    }
    _vars->registerVariable("end_iter", "unsigned int", "", U(std::numeric_limits<int32_t>::max()));
    _vars->registerVariable("end_frame", "unsigned int", "", U(std::numeric_limits<int32_t>::max()));
    _vars->registerVariable("N", "size_t", "", U(N));
    _vars->registerVariable("n_sets", "unsigned int", "", U(sd.sets.size()));
    _vars->registerVariable("n_radix", "size_t", "", U(n_radix));
    _vars->registerVariable("n_cells", "svec4", "", "1, 1, 1, 1");
    _vars->registerVariable("support", "float", "", "2");
    // default arrays (CalcServer.cpp:221-230)
    _vars->registerVariable("id", "size_t*", U(N), "");
    _vars->registerVariable("r", "vec*", U(N), "");
    _vars->registerVariable("iset", "unsigned int*", U(N), "");
    _vars->registerVariable("id_sorted", "size_t*", U(N), "");
    _vars->registerVariable("id_unsorted", "size_t*", U(N), "");
    _vars->registerVariable("icell", "size_t*", U(N), "");
    _vars->registerVariable("ihoc", "size_t*", "n_cells_w", "");
    // user variables, in document order
    for (auto& v : sd.variables)
        _vars->registerVariable(v.name, v.type, v.length, v.value);
    buildDefinitions();
    // tools
    for (auto& t : sd.tools) {
        Tool* tool = makeTool(*t);
        tool->id_in_pipeline((int)_tools.size());
        _tools.emplace_back(tool);
    }
    for (auto& r : sd.reports) {
        ProblemSetup::Tool t = *r;
        t.set("type", "report_" + r->get("type"));
        Tool* tool = makeTool(t);
        tool->id_in_pipeline((int)_tools.size());
        _tools.emplace_back(tool);
    }
    for (size_t i = 0; i + 1 < _tools.size(); i++)
        _tools[i]->next_tool(_tools[i + 1].get());
}

CalcServer::~CalcServer()
{
    _savers.clear(); // (joins the writer threads, frees their staging buffers)
    _tools.clear();
    if (_unsort_scratch)
        aqc_free(_ctx, _unsort_scratch);
    _vars.reset();
    if (_ctx)
        aqc_ctx_destroy(_ctx);
}

// CalcServer.cpp:240-265: evaluated definitions are solved as float and printed with
// "%#G" + "f"; the others are passed verbatim.  The CUDA kernels read the handful
// of definitions they know from the context (aqc_set_defs).
void CalcServer::buildDefinitions()
{
    for (auto& d : _sim_data.definitions) {
        std::string val = d.value;
        if (!d.value.empty() && d.evaluate) {
            float f = 0.f;
            _vars->solve("float", d.value, &f);
            char b[128];
            snprintf(b, sizeof(b), "%#G", f);
            val = std::string(b) + "f";
        }
        _defs.emplace_back(d.name, val);
    }
    auto lookup = [&](std::string n, int depth = 0) -> std::string {
        // resolve chains like __LAP_FORMULATION__ -> __LAP_MONAGHAN__ -> 1
        for (int k = 0; k < 8; k++) {
            bool found = false;
            for (auto& kv : _defs)
                if (kv.first == n) {
                    n = kv.second;
                    found = true;
                    break;
                }
            if (!found)
                break;
        }
        (void)depth;
        return n;
    };
    auto has = [&](const std::string& n) {
        for (auto& kv : _defs)
            if (kv.first == n)
                return true;
        return false;
    };
    auto num = [&](const std::string& n, float def) {
        if (!has(n))
            return def;
        std::string v = lookup(n);
        if (!v.empty() && (v.back() == 'f' || v.back() == 'F'))
            v.pop_back();
        return strtof(v.c_str(), nullptr);
    };
    aqc_defs D;
    D.dims = _sim_data.dims;
    D.H = num("H", 0.f);
    D.CONW = num("CONW", 0.f);
    D.CONF = num("CONF", 0.f);
    D.SUPPORT = num("SUPPORT", 2.f);
    D.DIMS = num("DIMS", (float)_sim_data.dims);
    if (has("H")) {
        if (aqc_set_defs(_ctx, &D))
            throw std::runtime_error(aqc_last_error(_ctx));
    }
    for (auto& kv : _defs) {
        const std::string val = lookup(kv.first);
        if (kv.first == "__LAP_FORMULATION__")
            _lap_morris = (val == "2" || val == "__LAP_MORRIS__");
        const int rc = aqc_set_define(_ctx, kv.first.c_str(), val.c_str());
        if (rc < 0)
            throw std::runtime_error(std::string("Unsupported definition: ") + aqc_last_error(_ctx));
    }
}

Tool* CalcServer::makeTool(const ProblemSetup::Tool& t)
{
    const std::string type = t.get("type"), name = t.get("name");
    const bool once = t.get("once") == "true";
    if (type == "kernel")
        return new Kernel(this, name, t.get("path"), t.get("entry_point"), t.get("n"), once);
    if (type == "copy")
        return new Copy(this, name, t.get("in"), t.get("out"), once);
    if (type == "set")
        return new Set(this, name, t.get("in"), t.get("value"), once);
    if (type == "set_scalar")
        return new SetScalar(this, name, t.get("in"), t.get("value"), once);
    if (type == "reduction")
        return new Reduction(this, name, t.get("in"), t.get("out"), t.get("operation"), t.get("null"), once);
    if (type == "link-list")
        return new LinkList(this, name, t, once);
    if (type == "radix-sort" || type == "sort")
        return new RadixSort(this, name, t.get("in"), t.get("perm"), t.get("inv_perm"), once);
    if (type == "unsort")
        return new UnSort(this, name, t.get("in"), t.get("out"), t.get("perm"), once);
    if (type == "assert")
        return new Assert(this, name, t.get("condition"), once);
    if (type == "if")
        return new If(this, name, t.get("condition"), once);
    if (type == "while")
        return new While(this, name, t.get("condition"), once);
    if (type == "endif" || type == "end")
        return new End(this, name, once);
    if (type == "mpi-sync")
        return new MPISync(this, name, t.get("mask"), t.get("fields"), t.get("processes"), t.get("depends"), once);
    if (type == "mpi-allreduce")
        return new MPIAllReduce(this, name, t.get("in"), t.get("operation"), once);
    if (type == "python")
        return new PythonTool(this, name, t.get("path"), once);
    if (type == "dummy")
        return new Dummy(this, name, once);
    if (type == "installable") {
        // CalcServer.cpp:375-400: dlopen(path), dlsym("create_object"), Tool* create_object(const
        // std::string name, bool once).  The plugin derives from THIS host's Tool
        // (aquagpusph_b200/host/calcserver.hpp) and links libaquahost.so, the way the reference's
        // tests/ExternalTool links libaquagpusphlib.so; a plugin built against the reference's
        // OpenCL Tool cannot be loaded (there is no OpenCL here).
        void* handle = dlopen(t.get("path").c_str(), RTLD_LAZY);
        if (!handle) {
            const char* why = dlerror(); // (a second call returns NULL: the message is consumed)
            throw std::runtime_error("Installable tool \"" + name + "\" failed loading \"" + t.get("path") +
                                     "\" library: " + (why ? why : "?"));
        }
        typedef Tool* (*maker_t)(const std::string, bool);
        maker_t maker = (maker_t)dlsym(handle, "create_object");
        if (!maker)
            throw std::runtime_error("Installable tool \"" + name +
                                     "\" failed loading \"create_object\" symbol");
        Tool* tool = maker(name, once);
        if (!tool)
            throw std::runtime_error("Installable tool \"" + name + "\": create_object returned NULL");
        tool->attach(this);
        return tool;
    }
    if (startswith(type, "report_"))
        return new Report(this, name, type.substr(7), t, once);
    throw std::runtime_error("The tool \"" + name + "\" has the type \"" + type +
                             "\", which this build does not provide");
}

void CalcServer::setup()
{
    Variable* h = _vars->get("h");
    if (!h)
        throw std::runtime_error("Undeclared kernel length variable \"h\"");
    if (h->type() != "float")
        throw std::runtime_error("Kernel length variable \"h\" must be of type \"float\"");
    if (*(float*)h->get() <= 0.f)
        throw std::runtime_error("Kernel length variable \"h\" must be positive");
    // per-set scalars (CalcServer.cpp:1466-1528)
    for (size_t i = 0; i < _sim_data.sets.size(); i++)
        for (auto& kv : _sim_data.sets[i]->scalars) {
            Variable* v = _vars->get(kv.first);
            if (!v)
                throw std::runtime_error("Particles set " + std::to_string(i) +
                                         " asks for the undeclared variable \"" + kv.first + "\"");
            if (!v->isArray())
                throw std::runtime_error("Particles set " + std::to_string(i) + ": \"" + kv.first +
                                         "\" must be an array");
            if (v->length() != _sim_data.sets.size())
                throw std::runtime_error("Particles set " + std::to_string(i) + ": \"" + kv.first +
                                         "\" must have length n_sets");
            std::vector<char> data(v->typesize());
            _vars->solve(v->type(), kv.second, data.data());
            if (aqc_memcpy_h2d(_ctx, (char*)v->dptr() + i * v->typesize(), data.data(), v->typesize(), 1))
                throw std::runtime_error(aqc_last_error(_ctx));
        }
    for (auto& t : _tools)
        t->setup();
    planFusion();
    planDeviceLoops();
}

// Sweep fusion planner.  A group {M0 < M1 < ...} of kernel tools is executed by ONE fused
// launch at M0's position when, for every member M and every tool T of the pipeline that
// sits between M0 and M (other members included):
//   * T's dependencies are known (kernel, copy; anything else ends the search),
//   * T does not write what M reads -- unless T is a kernel that only writes rows of a
//     particle class the fused sweep never reads (aqc_kernel_write_rows vs
//     aqc_fused_read_rows: cfd/Sensors.cl writes u, rho, p of the sensors only),
//   * T neither reads nor writes what M writes,
// and no member reads or writes another member's output.  Otherwise the tools run
// one by one exactly as listed.
void CalcServer::planFusion()
{
    if (getenv("AQUA_NO_FUSION") || _lap_morris) // (the fused groups hold the Monaghan build of cfd/Interactions.cl)
        return;
    auto has = [](const std::vector<Variable*>& v, Variable* x) {
        return std::find(v.begin(), v.end(), x) != v.end();
    };
    auto overlap = [&](const std::vector<Variable*>& a, const std::vector<Variable*>& b) {
        for (auto x : a)
            if (has(b, x))
                return true;
        return false;
    };
    for (size_t i = 0; i < _tools.size(); i++) {
        Kernel* lead = dynamic_cast<Kernel*>(_tools[i].get());
        if (!lead || lead->fused())
            continue;
        std::vector<Kernel*> members{ lead };
        std::vector<size_t> pos{ i };
        std::vector<int> ids{ lead->kernel_id() };
        if (aqc_fused_prefix(ids.data(), 1, dims()) <= 0)
            continue;
        int best_fid = -1;
        size_t best_n = 0;
        for (size_t j = i + 1; j < _tools.size() && j < i + 32; j++) {
            Tool* t = _tools[j].get();
            if (t->scope_modifier() != 0)
                break;
            Kernel* k = dynamic_cast<Kernel*>(t);
            if (k && !k->fused()) {
                ids.push_back(k->kernel_id());
                if (aqc_fused_prefix(ids.data(), (int)ids.size(), dims()) > 0) {
                    members.push_back(k);
                    pos.push_back(j);
                    const int fid = aqc_fused_lookup(ids.data(), (int)ids.size(), dims());
                    if (fid >= 0) {
                        best_fid = fid;
                        best_n = members.size();
                    }
                    continue;
                }
                ids.pop_back();
            }
            std::vector<Variable*> in, out;
            if (!t->dependencies(in, out))
                break;
        }
        if (best_fid < 0)
            continue;
        members.resize(best_n);
        pos.resize(best_n);
        // safety
        bool ok = true;
        const unsigned read_rows = (unsigned)aqc_fused_read_rows(best_fid);
        std::vector<std::vector<Variable*>> min(best_n), mout(best_n);
        for (size_t m = 0; m < best_n; m++)
            members[m]->dependencies(min[m], mout[m]);
        for (size_t m = 0; m < best_n && ok; m++) {
            for (size_t j = pos[0] + 1; j < pos[m] && ok; j++) {
                Tool* t = _tools[j].get();
                std::vector<Variable*> tin, tout;
                t->dependencies(tin, tout);
                const bool is_member = std::find(pos.begin(), pos.end(), j) != pos.end();
                if (is_member) {
                    // an earlier member must not produce what M consumes, nor share outputs
                    if (overlap(tout, min[m]) || overlap(tout, mout[m]) || overlap(tin, mout[m]))
                        ok = false;
                    continue;
                }
                Kernel* tk = dynamic_cast<Kernel*>(t);
                const unsigned wr = tk ? (unsigned)aqc_kernel_write_rows(tk->kernel_id()) : 7u;
                if (overlap(tout, min[m]) && (wr & read_rows))
                    ok = false;
                if (overlap(tin, mout[m]) || overlap(tout, mout[m]))
                    ok = false;
            }
        }
        if (!ok)
            continue;
        lead->fuse_lead(best_fid, members);
        for (size_t m = 1; m < best_n; m++)
            members[m]->fuse_follow(lead);
        _fused_groups++;
        if (logLevel() <= 1) {
            std::string msg = "Fused sweep:";
            for (auto k : members)
                msg += " \"" + k->name() + "\"";
            fprintf(stderr, "INFO: %s\n", msg.c_str());
        }
    }
}

void CalcServer::commInit(const void* unique_id)
{
    if (aqc_comm_init(_ctx, _mpi_rank, _mpi_size, unique_id))
        throw std::runtime_error(std::string("Cannot join the communicator: ") + aqc_last_error(_ctx));
}

void CalcServer::step()
{
    Tool* tool = _tools.empty() ? nullptr : _tools.front().get();
    while (tool) {
        tool->execute();
        tool = tool->next_tool();
    }
    _steps++;
}

void CalcServer::update(TimeManager& t)
{
    while (!t.mustPrintOutput() && !t.mustStop())
        step();
}

void CalcServer::download(const std::string& var, void* out)
{
    Variable* v = _vars->get(var);
    if (!v || !v->isArray())
        throw std::runtime_error("No such array \"" + var + "\"");
    if (aqc_memcpy_d2h(_ctx, out, v->dptr(), v->size(), 1))
        throw std::runtime_error(aqc_last_error(_ctx));
}

void CalcServer::upload(const std::string& var, const void* in)
{
    Variable* v = _vars->get(var);
    if (!v || !v->isArray())
        throw std::runtime_error("No such array \"" + var + "\"");
    if (aqc_memcpy_h2d(_ctx, v->dptr(), in, v->size(), 1))
        throw std::runtime_error(aqc_last_error(_ctx));
}

void CalcServer::getUnsortedMem(const std::string& var, void* out)
{
    Variable* v = _vars->get(var);
    Variable* id = _vars->get("id");
    if (!v || !v->isArray())
        throw std::runtime_error("No such array \"" + var + "\"");
    if (v->length() != id->length()) { // not a per-particle array
        download(var, out);
        return;
    }
    if (v->size() > _unsort_cap) {
        if (_unsort_scratch)
            aqc_free(_ctx, _unsort_scratch);
        if (aqc_alloc(_ctx, v->size(), &_unsort_scratch))
            throw std::runtime_error(aqc_last_error(_ctx));
        _unsort_cap = v->size();
    }
    const void* src[1] = { v->dptr() };
    void* dst[1] = { _unsort_scratch };
    const size_t eb[1] = { v->typesize() };
    if (aqc_scatter_fields(_ctx, (const aqc_usize*)id->dptr(), (aqc_usize)v->length(), 1, src, dst, eb) ||
        aqc_memcpy_d2h(_ctx, out, _unsort_scratch, v->size(), 1))
        throw std::runtime_error(aqc_last_error(_ctx));
}

// ------------------------------------------------------------ particles I/O --
static bool isSep(char c) { return isspace((unsigned char)c) || strchr(",;()[]{}", c); }

void CalcServer::loadParticles()
{
    Variables* vars = _vars.get();
    size_t offset = 0;
    for (size_t iset = 0; iset < _sim_data.sets.size(); iset++) {
        auto& set = *_sim_data.sets[iset];
        const size_t n = set.n;
        // Particles::loadDefault (Particles.cpp:122-223)
        {
            std::vector<uint32_t> is(n, (uint32_t)iset), id(n);
            for (size_t i = 0; i < n; i++)
                id[i] = (uint32_t)(offset + i);
            auto put = [&](const char* name, const void* src) {
                Variable* v = vars->get(name);
                if (aqc_memcpy_h2d(_ctx, (char*)v->dptr() + offset * 4, src, n * 4, 1))
                    throw std::runtime_error(aqc_last_error(_ctx));
            };
            put("iset", is.data());
            put("id", id.data());
            put("id_sorted", id.data());
            put("id_unsorted", id.data());
        }
        if (!set.in_path.empty()) {
            const std::string fmt = toLowerCopy(set.in_format);
            if (fmt != "fastascii" && fmt != "ascii" && fmt != "csv")
                throw std::runtime_error("Particles set " + std::to_string(iset) +
                                         ": unsupported input format \"" + set.in_format + "\"");
            std::vector<Variable*> fields;
            for (auto f : split(set.in_fields)) {
                Variable* v = vars->get(f);
                if (!v || !v->isArray())
                    throw std::runtime_error("Particles set " + std::to_string(iset) +
                                             ": undeclared field \"" + f + "\"");
                fields.push_back(v);
            }
            std::vector<std::vector<char>> host;
            for (auto v : fields)
                host.emplace_back(n * v->typesize());
            const std::string path = formatPath(set.in_path, _mpi_rank);
            std::ifstream in(path);
            if (!in)
                throw std::runtime_error("Particles set " + std::to_string(iset) + ": cannot read \"" +
                                         path + "\"");
            std::string line;
            size_t i = 0;
            while (i < n && std::getline(in, line)) {
                const char* p = line.c_str();
                while (*p && isspace((unsigned char)*p))
                    p++;
                if (!*p || *p == '#')
                    continue;
                for (size_t k = 0; k < fields.size(); k++) {
                    Variable* v = fields[k];
                    char* dst = host[k].data() + i * v->typesize();
                    for (unsigned c = 0; c < v->ncomp(); c++) {
                        while (*p && isSep(*p))
                            p++;
                        if (!*p)
                            throw std::runtime_error("Particles set " + std::to_string(iset) + ", \"" +
                                                     path + "\": not enough fields in line \"" + line + "\"");
                        char* end;
                        switch (v->kind()) {
                            case 'u': { uint32_t x = (uint32_t)strtoul(p, &end, 10); memcpy(dst + 4 * c, &x, 4); break; }
                            case 'i': { int32_t x = (int32_t)strtol(p, &end, 10); memcpy(dst + 4 * c, &x, 4); break; }
                            default: { float x = strtof(p, &end); memcpy(dst + 4 * c, &x, 4); }
                        }
                        if (end == p)
                            throw std::runtime_error("Particles set " + std::to_string(iset) + ", \"" +
                                                     path + "\": cannot parse \"" + line + "\"");
                        p = end;
                    }
                }
                i++;
            }
            if (i != n)
                throw std::runtime_error("Particles set " + std::to_string(iset) + ": \"" + path +
                                         "\" holds " + std::to_string(i) + " particles, " +
                                         std::to_string(n) + " expected");
            for (size_t k = 0; k < fields.size(); k++)
                if (aqc_memcpy_h2d(_ctx, (char*)fields[k]->dptr() + offset * fields[k]->typesize(),
                                   host[k].data(), host[k].size(), 1))
                    throw std::runtime_error(aqc_last_error(_ctx));
        }
        offset += n;
    }
}

} // namespace CalcServer
} // namespace Aqua
