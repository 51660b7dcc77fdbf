// devloop.cpp -- see devloop.hpp
#include "devloop.hpp"

#include <algorithm>
#include <cstring>

#include "calcserver.hpp"

namespace Aqua {
namespace CalcServer {

using InputOutput::Variable;
using InputOutput::Variables;

// ------------------------------------------------------------- SvmCompiler --
namespace {

aqs_op mk(int code, int a = 0, int b = 0, int c = 0, double imm = 0.0)
{
    aqs_op o;
    o.code = code;
    o.a = a;
    o.b = b;
    o.c = c;
    o.imm = imm;
    return o;
}

// The productions of Tokenizer::P (host/tokenizer.hpp), emitting instead of evaluating
struct Emitter {
    const std::string& s;
    size_t i;
    const Tokenizer* tok;
    SvmCompiler::Resolver resolve;
    void* user;
    std::vector<aqs_op>& out;

    void skip()
    {
        while (i < s.size() && isspace((unsigned char)s[i]))
            i++;
    }
    bool eat(const char* t)
    {
        skip();
        const size_t n = strlen(t);
        if (s.compare(i, n, t))
            return false;
        i += n;
        return true;
    }
    [[noreturn]] void fail(const std::string& why)
    {
        throw std::runtime_error("Error compiling \"" + s + "\": " + why + " at position " + std::to_string(i));
    }
    void ternary()
    {
        logic_or();
        skip();
        if (i < s.size() && s[i] == '?') {
            i++;
            ternary();
            if (!eat(":"))
                fail("':' expected");
            ternary();
            out.push_back(mk(AQS_SELECT));
        }
    }
    void logic_or()
    {
        logic_and();
        while (eat("||")) {
            logic_and();
            out.push_back(mk(AQS_OR));
        }
    }
    void logic_and()
    {
        compare();
        while (eat("&&")) {
            compare();
            out.push_back(mk(AQS_AND));
        }
    }
    void compare()
    {
        additive();
        for (;;) {
            int code;
            if (eat("<=")) code = AQS_LE;
            else if (eat(">=")) code = AQS_GE;
            else if (eat("==")) code = AQS_EQ;
            else if (eat("!=")) code = AQS_NE;
            else if (eat("<")) code = AQS_LT;
            else if (eat(">")) code = AQS_GT;
            else return;
            additive();
            out.push_back(mk(code));
        }
    }
    void additive()
    {
        term();
        for (;;) {
            int code;
            if (eat("+")) code = AQS_ADD;
            else if (eat("-")) code = AQS_SUB;
            else return;
            term();
            out.push_back(mk(code));
        }
    }
    void term()
    {
        unary();
        for (;;) {
            int code;
            if (eat("*")) code = AQS_MUL;
            else if (eat("/")) code = AQS_DIV;
            else if (eat("%")) code = AQS_MOD;
            else return;
            unary();
            out.push_back(mk(code));
        }
    }
    void unary()
    {
        skip();
        if (i < s.size() && s[i] == '-') {
            i++;
            unary();
            out.push_back(mk(AQS_NEG));
            return;
        }
        if (i < s.size() && s[i] == '+') {
            i++;
            unary();
            return;
        }
        if (i < s.size() && s[i] == '!' && (i + 1 >= s.size() || s[i + 1] != '=')) {
            i++;
            unary();
            out.push_back(mk(AQS_NOT));
            return;
        }
        power();
    }
    void power()
    {
        primary();
        skip();
        if (i < s.size() && s[i] == '^') {
            i++;
            unary(); // right associative
            out.push_back(mk(AQS_POW));
        }
    }
    void primary()
    {
        skip();
        if (i >= s.size())
            fail("unexpected end");
        if (s[i] == '(') {
            i++;
            ternary();
            if (!eat(")"))
                fail("')' expected");
            return;
        }
        if (isdigit((unsigned char)s[i]) || s[i] == '.') {
            char* end;
            const double v = strtod(s.c_str() + i, &end);
            const size_t j = end - s.c_str();
            if (j == i)
                fail("bad number");
            i = j;
            if (i < s.size() && (s[i] == 'f' || s[i] == 'F'))
                i++;
            out.push_back(mk(AQS_IMM, 0, 0, 0, v));
            return;
        }
        if (isalpha((unsigned char)s[i]) || s[i] == '_') {
            size_t j = i;
            while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_'))
                j++;
            const std::string id = s.substr(i, j - i);
            i = j;
            skip();
            if (i < s.size() && s[i] == '(') {
                i++;
                int n = 0;
                skip();
                if (i < s.size() && s[i] == ')') {
                    i++;
                } else {
                    for (;;) {
                        ternary();
                        n++;
                        if (eat(","))
                            continue;
                        if (eat(")"))
                            break;
                        fail("',' or ')' expected");
                    }
                }
                call(id, n);
                return;
            }
            SvmCompiler::Slot slot;
            if (resolve && resolve(user, id, slot)) {
                out.push_back(mk(AQS_LOAD, slot.offset, slot.kind));
                return;
            }
            if (!tok->isVariable(id))
                fail("unknown variable \"" + id + "\"");
            out.push_back(mk(AQS_IMM, 0, 0, 0, tok->variable(id)));
            return;
        }
        fail(std::string("unexpected character '") + s[i] + "'");
    }
    void call(const std::string& f, int n)
    {
        static const struct {
            const char* name;
            int id, nargs;
        } fns[] = { { "sqrt", AQS_F_SQRT, 1 },   { "abs", AQS_F_ABS, 1 },     { "sin", AQS_F_SIN, 1 },
                    { "cos", AQS_F_COS, 1 },     { "tan", AQS_F_TAN, 1 },     { "asin", AQS_F_ASIN, 1 },
                    { "acos", AQS_F_ACOS, 1 },   { "atan", AQS_F_ATAN, 1 },   { "atan2", AQS_F_ATAN2, 2 },
                    { "sinh", AQS_F_SINH, 1 },   { "cosh", AQS_F_COSH, 1 },   { "tanh", AQS_F_TANH, 1 },
                    { "exp", AQS_F_EXP, 1 },     { "log", AQS_F_LOG, 1 },     { "ln", AQS_F_LOG, 1 },
                    { "log2", AQS_F_LOG2, 1 },   { "log10", AQS_F_LOG10, 1 }, { "floor", AQS_F_FLOOR, 1 },
                    { "ceil", AQS_F_CEIL, 1 },   { "round", AQS_F_RINT, 1 },  { "rint", AQS_F_RINT, 1 },
                    { "sign", AQS_F_SIGN, 1 },   { "min", AQS_F_MIN, -1 },    { "max", AQS_F_MAX, -1 },
                    { "sum", AQS_F_SUM, -1 },    { "avg", AQS_F_AVG, -1 } };
        if (f == "pow") {
            if (n != 2)
                fail("function pow expects 2 arguments");
            out.push_back(mk(AQS_POW));
            return;
        }
        for (auto& e : fns)
            if (f == e.name) {
                if (e.nargs >= 0 && n != e.nargs)
                    fail("function " + f + " expects " + std::to_string(e.nargs) + " arguments");
                if (e.nargs < 0 && n == 0)
                    fail("function " + f + " needs arguments");
                if (n > AQS_STACK / 2)
                    fail("function " + f + ": too many arguments");
                out.push_back(mk(AQS_CALL, e.id, n));
                return;
            }
        fail("unknown function \"" + f + "\"");
    }
};

} // namespace

void SvmCompiler::compile(const std::string& expr, std::vector<aqs_op>& out) const
{
    Emitter p{ expr, 0, _tok, _resolve, _user, out };
    p.skip();
    if (p.i >= expr.size())
        throw std::runtime_error("Empty expression");
    p.ternary();
    p.skip();
    if (p.i != expr.size())
        throw std::runtime_error("Error compiling \"" + expr + "\": unexpected token at " +
                                 std::to_string(p.i));
}

// -------------------------------------------------------------- DeviceLoop --
DeviceLoop::DeviceLoop(CalcServer* C, Tool* opener, const std::string& condition, size_t first, size_t last)
  : _C(C), _opener(opener), _condition(condition), _first(first), _last(last)
{
    if (const char* e = getenv("AQUA_DEVLOOP_MAX_ITERS"))
        _max_iters = (uint32_t)std::max(1l, atol(e));
}

DeviceLoop::~DeviceLoop()
{
    if (_loop)
        aqc_loop_destroy(_C->ctx(), _loop);
    for (void* e : _events)
        if (e)
            aqc_event_destroy(_C->ctx(), e);
    if (_ev_fork)
        aqc_event_destroy(_C->ctx(), _ev_fork);
}

unsigned DeviceLoop::branchTools() const
{
    unsigned n = 0;
    for (int l : _lane_of)
        n += l == 1;
    return _two_lanes ? n : 0;
}

// The schedule itself, free of tools and devices (tests/test_host_cpu.py checks it on random dependency
// sets through aqh_lane_schedule: every conflicting pair ends up ordered).  List scheduling in pipeline
// order: a tool that launches something goes to lane 1 when it could start there at least `gain` cost units
// earlier than on lane 0; then, for every tool, the LAST conflicting tool of the other lane is the one
// whose event it waits for (the lanes are in-order queues: that covers the earlier ones), unless its lane
// has already waited for a later one.
void scheduleLanes(const std::vector<LaneDep>& deps, double gain, std::vector<int>& lane_of,
                   std::vector<std::vector<int>>& waits, std::vector<char>& marked, int& last_lane1)
{
    const size_t n = deps.size();
    lane_of.assign(n, 0);
    waits.assign(n, {});
    marked.assign(n, 0);
    last_lane1 = -1;
    auto meets = [](const LaneDep::Access& a, const LaneDep::Access& b) {
        for (auto& x : a)
            for (auto& y : b)
                if (x.first == y.first && (x.second & y.second))
                    return true;
        return false;
    };
    auto conflict = [&](const LaneDep& a, const LaneDep& b) {
        return a.barrier || b.barrier || meets(a.w, b.r) || meets(a.w, b.w) || meets(a.r, b.w);
    };
    double free_at[2] = { 0.0, 0.0 };
    std::vector<double> finish(n, 0.0);
    for (size_t k = 0; k < n; k++) {
        const LaneDep& d = deps[k];
        if (!d.launches)
            continue;
        double ready = 0.0;
        for (size_t u = 0; u < k; u++)
            if (deps[u].launches && conflict(deps[u], d))
                ready = std::max(ready, finish[u]);
        const double r0 = std::max(ready, free_at[0]), r1 = std::max(ready, free_at[1]);
        const int l = (!d.forced0 && !d.barrier && r1 + gain <= r0) ? 1 : 0;
        lane_of[k] = l;
        finish[k] = (l ? r1 : r0) + d.cost;
        free_at[l] = finish[k];
    }
    int waited[2] = { -1, -1 }; // per lane: the newest tool of the OTHER lane it has waited for
    for (size_t k = 0; k < n; k++) {
        if (!deps[k].launches)
            continue;
        const int l = lane_of[k];
        int last = -1;
        for (size_t u = 0; u < k; u++)
            if (deps[u].launches && lane_of[u] != l && conflict(deps[u], deps[k]))
                last = (int)u;
        if (last > waited[l]) {
            waits[k].push_back(last);
            marked[last] = 1;
            waited[l] = last;
        }
        if (l == 1)
            last_lane1 = (int)k;
    }
    if (last_lane1 >= 0)
        marked[last_lane1] = 1; // the join at the end of the pass
}

// Which lane every tool of the body runs on, and the events it waits for.  Dependencies are whole
// arrays (Tool::dependencies: what the fusion planner uses too): two tools conflict when one writes
// what the other reads or writes.  Lanes come from list scheduling in pipeline order with a crude
// cost (a neighbour sweep = 10 element-wise kernels): a tool goes to the branch lane when it could
// start there clearly earlier.  Whatever the assignment, every conflicting pair on different lanes
// is ordered by an event, so the result is the one of the pipeline order.
void DeviceLoop::planLanes()
{
    auto& tools = _C->tools();
    const size_t n = _last - _first;
    _lane_of.assign(n, 0);
    _waits.assign(n, {});
    _marked.assign(n, 0);
    _events.assign(n, nullptr);
    _two_lanes = false;
    _last_lane1 = -1;
    if (const char* e = getenv("AQUA_DEVICE_LANES"))
        if (!strcmp(e, "0"))
            return;
    // what a tool reads / writes: arrays with the particle classes (AQC_ROWS_*) of the rows involved
    // -- cfd/Sensors.cl writes sensor rows only, the fused fluid sweep reads fluid rows only (what
    // the fusion planner relies on as well), BIe interactions boundary rows ...
    typedef std::vector<std::pair<const Variable*, unsigned>> Access;
    struct Dep {
        Access r, w;
        bool barrier = false, forced0 = false, launches = true;
        double cost = 1.0;
    };
    std::vector<Dep> deps(n);
    auto add = [](Access& a, const Variable* v, unsigned rows) {
        if (!v || !v->isArray())
            return;
        for (auto& e : a)
            if (e.first == v) {
                e.second |= rows;
                return;
            }
        a.emplace_back(v, rows);
    };
    auto add_all = [&](Access& a, const std::vector<Variable*>& vs, unsigned rows) {
        for (auto v : vs)
            add(a, v, rows);
    };
    // (off by default: with the row classes the sensors, the BIe sets and BIe interactions move next
    // to the fused fluid sweep, which loses more than they gain -- 25.14 against 24.87 ms per step at
    // 1.2 M particles, 7.52 against 6.97 at 142 k; profiles/r2_lanes_assignment.txt)
    const bool use_rows = getenv("AQUA_LANE_ROWS") && !strcmp(getenv("AQUA_LANE_ROWS"), "1");
    for (size_t k = 0; k < n; k++) {
        Tool* t = tools[_first + k].get();
        Dep& d = deps[k];
        std::vector<Variable*> in, out;
        if (Kernel* kt = dynamic_cast<Kernel*>(t)) {
            if (kt->leader()) { // runs inside its leader's launch
                d.launches = false;
                d.cost = 0.0;
                continue;
            }
            std::vector<Kernel*> members = kt->group();
            const bool fused = !members.empty();
            if (!fused)
                members.push_back(kt);
            d.cost = 0.0;
            for (auto m : members) {
                in.clear();
                out.clear();
                m->dependencies(in, out);
                d.cost += m->isSweep() ? 10.0 : 1.0;
                unsigned rrows = AQC_ROWS_ANY, wrows = AQC_ROWS_ANY;
                if (use_rows) {
                    rrows = fused ? (unsigned)aqc_fused_read_rows(kt->fusedId())
                                  : (unsigned)aqc_kernel_read_rows(m->kernel_id());
                    wrows = fused ? (unsigned)AQC_ROWS_ANY : (unsigned)aqc_kernel_write_rows(m->kernel_id());
                }
                for (auto v : in) // (positions are read whatever the class)
                    add(d.r, v, (v->name() == "r" || v->name() == "r_in") ? (unsigned)AQC_ROWS_ANY : rrows);
                add_all(d.w, out, wrows);
                add_all(d.r, out, AQC_ROWS_ANY); // (an output may be read back: accumulated into, or p[j] of p_boundary)
                for (auto v : m->arguments())
                    if (!v->isArray() && varying(v))
                        d.forced0 = true; // reads the table: behind the scalar programs
            }
            if (use_rows && !fused && aqc_kernel_read_rows(kt->kernel_id()) != AQC_ROWS_ANY)
                for (auto& e : d.r) // a restricted reader reads its outputs' other rows the same way
                    for (auto v : out)
                        if (e.first == v)
                            e.second = (unsigned)aqc_kernel_read_rows(kt->kernel_id()) |
                                       (unsigned)aqc_kernel_write_rows(kt->kernel_id());
            continue;
        } else if (dynamic_cast<Reduction*>(t)) {
            t->dependencies(in, out);
            d.forced0 = true; // (the reduction scratch and the table are lane 0's)
        } else if (dynamic_cast<Copy*>(t) || dynamic_cast<Set*>(t)) {
            t->dependencies(in, out);
        } else if (dynamic_cast<ScalarExpression*>(t) || dynamic_cast<Report*>(t) || dynamic_cast<Dummy*>(t)) {
            d.launches = false; // scalar programs: queued, run on lane 0
            d.forced0 = true;
            d.cost = 0.0;
            continue;
        } else {
            d.barrier = true;
            d.forced0 = true;
        }
        add_all(d.r, in, AQC_ROWS_ANY);
        add_all(d.w, out, AQC_ROWS_ANY);
        add_all(d.r, out, AQC_ROWS_ANY);
    }
    // (tuning knobs; the defaults are what was measured best on the dam break at 1.2 M particles)
    double gain = 3.0, sweep_cost = 10.0;
    if (const char* e = getenv("AQUA_LANE_GAIN"))
        gain = atof(e);
    if (const char* e = getenv("AQUA_LANE_SWEEP_COST"))
        sweep_cost = atof(e);
    std::vector<LaneDep> ld(n);
    for (size_t k = 0; k < n; k++) {
        for (auto& e : deps[k].r)
            ld[k].r.emplace_back((const void*)e.first, e.second);
        for (auto& e : deps[k].w)
            ld[k].w.emplace_back((const void*)e.first, e.second);
        ld[k].barrier = deps[k].barrier;
        ld[k].forced0 = deps[k].forced0;
        ld[k].launches = deps[k].launches;
        ld[k].cost = deps[k].cost >= 10.0 ? deps[k].cost / 10.0 * sweep_cost : deps[k].cost;
    }
    scheduleLanes(ld, gain, _lane_of, _waits, _marked, _last_lane1);
    _two_lanes = _last_lane1 >= 0;
    if (_two_lanes && logLevel() <= L_INFO) {
        std::string msg = "The loop \"" + _opener->name() + "\" runs these tools on a second stream:";
        for (size_t k = 0; k < n; k++)
            if (deps[k].launches && _lane_of[k] == 1)
                msg += " \"" + tools[_first + k]->name() + "\"";
        log(L_INFO, msg + "\n");
    }
}

void DeviceLoop::lane(int l)
{
    if (l == _cur_lane)
        return;
    if (aqc_lane_select(_C->ctx(), l))
        throw std::runtime_error(aqc_last_error(_C->ctx()));
    _cur_lane = l;
}

aqc_ctx* DeviceLoop::ctx() const { return _C->ctx(); }

bool DeviceLoop::contains(const Tool* t) const
{
    const int i = t ? t->id_in_pipeline() : -1;
    return i >= (int)_first && i < (int)_last && _C->tools()[i].get() == t;
}

int DeviceLoop::offset(const Variable* v) const
{
    auto it = _slots.find(v);
    return it == _slots.end() ? -1 : it->second;
}

const void* DeviceLoop::deviceAddress(const Variable* v) const
{
    const int off = offset(v);
    return (off < 0 || !_loop) ? nullptr : (const char*)aqc_loop_table(_loop) + off;
}

void* DeviceLoop::scratchDevice(int off) const { return _loop ? (char*)aqc_loop_table(_loop) + off : nullptr; }

int DeviceLoop::scratch(const Tool* t, int which)
{
    auto key = std::make_pair(t, which);
    auto it = _scratch.find(key);
    if (it != _scratch.end())
        return it->second;
    if (_planned)
        throw std::runtime_error("DeviceLoop::scratch asked after the table was laid out");
    const int off = _table_bytes;
    _table_bytes += 16;
    _scratch[key] = off;
    return off;
}

void DeviceLoop::setInitial(int off, const void* data, size_t bytes)
{
    if (_initial.size() < (size_t)_table_bytes)
        _initial.resize(_table_bytes, 0);
    memcpy(_initial.data() + off, data, bytes);
}

bool DeviceLoop::resolve(void* user, const std::string& id, SvmCompiler::Slot& out)
{
    const DeviceLoop* L = (const DeviceLoop*)user;
    Variables* vars = L->_C->variables();
    Variable* v = vars->get(id);
    unsigned comp = 0;
    if (!v || v->isArray()) {
        static const char* suf[] = { "_x", "_y", "_z", "_w" };
        v = nullptr;
        for (unsigned c = 0; c < 4; c++)
            if (endswith(id, suf[c])) {
                Variable* b = vars->get(id.substr(0, id.size() - 2));
                if (b && !b->isArray() && b->ncomp() > c) {
                    v = b;
                    comp = c;
                }
                break;
            }
        if (!v)
            return false;
    } else if (v->ncomp() != 1) {
        return false; // a vector is only visible through its components (Variables::populate)
    }
    const int off = L->offset(v);
    if (off < 0)
        return false;
    out.offset = off + 4 * (int)comp;
    out.kind = v->kind();
    return true;
}

void DeviceLoop::compile(const std::string& expr, std::vector<aqs_op>& out) const
{
    SvmCompiler(&_C->variables()->tokenizer(), &DeviceLoop::resolve, (void*)this).compile(expr, out);
}

bool DeviceLoop::compilable(const std::string& expr, std::string& why) const
{
    try {
        for (auto& part : split_formulae(expr)) {
            std::vector<aqs_op> tmp;
            compile(part, tmp);
        }
    } catch (std::exception& e) {
        why = e.what();
        return false;
    }
    return true;
}

bool DeviceLoop::readsVarying(const std::string& expr) const
{
    for (auto v : _C->variables()->exprVariables(expr))
        if (varying(v))
            return true;
    return false;
}

bool DeviceLoop::plan(std::string& why)
{
    _usable = false;
    auto& tools = _C->tools();
    if (_first >= _last) {
        why = "empty body";
        return false;
    }
    // the table: every scalar the body writes
    for (size_t i = _first; i < _last; i++) {
        Tool* t = tools[i].get();
        if (t->once()) {
            why = "the tool \"" + t->name() + "\" runs once";
            return false;
        }
        std::vector<Variable*> outs;
        t->scalarOutputs(outs);
        for (auto v : outs) {
            if (!v || v->isArray() || _slots.count(v))
                continue;
            if ((v->kind() != 'f' && v->kind() != 'u' && v->kind() != 'i') || v->typesize() > 16) {
                why = "the tool \"" + t->name() + "\" writes \"" + v->name() + "\" of type \"" + v->type() +
                      "\" (only 32-bit scalars and vectors of up to 4 live on the device)";
                return false;
            }
            _slots[v] = _table_bytes;
            _order.push_back(v);
            _table_bytes += 16;
        }
    }
    if (_slots.empty()) {
        why = "the body writes no scalar: its condition cannot change";
        return false;
    }
    unsigned reports = 0;
    for (size_t i = _first; i < _last; i++) {
        Tool* t = tools[i].get();
        std::string w;
        if (!t->recordable(*this, w)) {
            why = "the tool \"" + t->name() + "\": " + w;
            return false;
        }
        if (dynamic_cast<Report*>(t))
            reports++;
    }
    std::string w;
    if (!compilable(_condition, w)) {
        why = "the condition: " + w;
        return false;
    }
    if (!readsVarying(_condition)) {
        why = "the condition reads nothing the body writes";
        return false;
    }
    _planned = true;
    if (_initial.size() < (size_t)_table_bytes)
        _initial.resize(_table_bytes, 0);
    // report snapshots kept per run: a loop that reports more often than this prints the rest as
    // "(not kept)" -- the midpoint loops end after 2 .. ~30 passes
    _hist_rows = reports ? (int)std::min<size_t>(4096, 256 * (size_t)reports) : 0;
    while (_hist_rows > 16 && (size_t)_hist_rows * (16 + (size_t)_table_bytes) > (4u << 20))
        _hist_rows /= 2;
    if (aqc_loop_create(_C->ctx(), _table_bytes, _hist_rows, 8192, &_loop)) {
        why = aqc_last_error(_C->ctx());
        return false;
    }
    planLanes();
    _usable = true;
    return true;
}

void DeviceLoop::flush()
{
    if (_pending.empty())
        return;
    if (_cur_lane != 0)
        throw std::runtime_error("DeviceLoop::flush on the branch lane");
    if (aqc_loop_svm(_C->ctx(), _loop, _pending.data(), (int)_pending.size()))
        throw std::runtime_error(aqc_last_error(_C->ctx()));
    _pending.clear();
}

void DeviceLoop::pass()
{
    // one pass over the body through the tools' record(): queued while the loop records, run at
    // once (in stream order, nothing read back) otherwise
    auto& tools = _C->tools();
    std::vector<aqs_op> cond;
    compile(_condition, cond);
    cond.push_back(mk(AQS_SETCOND));
    _pending.clear();
    auto chk = [&](int rc) {
        if (rc)
            throw std::runtime_error(aqc_last_error(_C->ctx()));
    };
    if (!_two_lanes) {
        for (size_t i = _first; i < _last; i++)
            tools[i]->record(*this);
    } else {
        try {
            // fork: the branch lane starts behind everything queued so far (the pass before this one)
            lane(0);
            chk(aqc_lane_event(_C->ctx(), &_ev_fork));
            lane(1);
            chk(aqc_lane_wait(_C->ctx(), _ev_fork));
            for (size_t i = _first; i < _last; i++) {
                const size_t k = i - _first;
                lane(_lane_of[k]);
                for (int u : _waits[k])
                    chk(aqc_lane_wait(_C->ctx(), _events[u]));
                tools[i]->record(*this);
                if (_marked[k])
                    chk(aqc_lane_event(_C->ctx(), &_events[k]));
            }
            // join: lane 0 goes on behind the branch
            lane(0);
            if (_last_lane1 >= 0)
                chk(aqc_lane_wait(_C->ctx(), _events[_last_lane1]));
        } catch (...) {
            try {
                lane(0);
            } catch (...) {
            }
            throw;
        }
    }
    emit(cond);
    flush();
}

void DeviceLoop::record()
{
    // the graph starts from the condition the pass before it left in the header
    const aqs_op entry = mk(AQS_RECOND);
    if (aqc_loop_begin(_C->ctx(), _loop, &entry, 1))
        throw std::runtime_error(aqc_last_error(_C->ctx()));
    try {
        pass();
    } catch (...) {
        aqc_loop_abort(_C->ctx(), _loop);
        throw;
    }
    if (aqc_loop_end(_C->ctx(), _loop))
        throw std::runtime_error(aqc_last_error(_C->ctx()));
}

bool DeviceLoop::run()
{
    if (!_usable)
        return false;
    auto& tools = _C->tools();
    Variables* vars = _C->variables();
    std::vector<char> table(_initial);
    for (auto v : _order)
        memcpy(table.data() + _slots[v], v->get(), v->typesize());
    if (aqc_loop_start(_C->ctx(), _loop, table.data(), _max_iters))
        throw std::runtime_error(aqc_last_error(_C->ctx()));
    // the first pass, tool by tool: it builds the neighbour lists and sizes every scratch buffer.
    // Its reductions and scalar tools already work on the device table, so nothing is read back
    // and the device is still busy with it while the body is recorded and instantiated below
    pass();
    bool recorded = true;
    try {
        record();
    } catch (std::exception& e) {
        // nothing of the recording ran; the caches it touched were invalidated by the library
        recorded = false;
        _failures++;
        log(_failures == 1 ? L_WARNING : L_DEBUG,
            "The loop \"" + _opener->name() + "\" could not be recorded (" + e.what() +
                "): it goes on tool by tool\n");
        if (_failures >= 3) {
            log(L_WARNING, "The loop \"" + _opener->name() + "\" stays on the host from now on\n");
            _usable = false;
        }
    }
    if (recorded)
        _failures = 0;
    aqs_header hdr;
    std::vector<char> hist((size_t)_hist_rows * (16 + (size_t)_table_bytes));
    if (aqc_loop_run(_C->ctx(), _loop, nullptr, _max_iters, &hdr, table.data(),
                     hist.empty() ? nullptr : hist.data()))
        throw std::runtime_error("Failure running the loop \"" + _opener->name() +
                                 "\" on the device: " + aqc_last_error(_C->ctx()));
    const unsigned passes = 1 + (recorded ? hdr.iters : 0);
    _runs++;
    _iterations += passes;
    // reports, in the order they happened, each seeing the scalars of its moment
    auto load = [&](const char* tab) {
        for (auto v : _order) {
            v->set(tab + _slots[v]);
            vars->populate(v);
        }
    };
    const unsigned kept = std::min<unsigned>(hdr.snaps, (unsigned)_hist_rows);
    for (unsigned k = 0; k < kept; k++) {
        const char* row = hist.data() + (size_t)k * (16 + (size_t)_table_bytes);
        int id;
        memcpy(&id, row, 4);
        if (id < (int)_first || id >= (int)_last)
            continue;
        load(row + 16);
        tools[id]->execute();
    }
    if (hdr.snaps > kept)
        log(L_WARNING, "The loop \"" + _opener->name() + "\": " + std::to_string(hdr.snaps - kept) +
                           " report lines were not kept\n");
    load(table.data());
    for (size_t i = _first; i < _last; i++)
        if (!dynamic_cast<Report*>(tools[i].get()))
            tools[i]->account(passes);
    if (hdr.error) {
        const std::string where = "the loop \"" + _opener->name() + "\" (device side)";
        if (hdr.error >= 0x40000u)
            throw std::runtime_error("A condition of " + where + " overflows an int");
        if (hdr.error >= 0x30000u)
            throw std::runtime_error(where + " did not end within " + std::to_string(_max_iters) +
                                     " iterations (AQUA_DEVLOOP_MAX_ITERS)");
        if (hdr.error >= 0x20000u)
            throw std::runtime_error("An expression of " + where + " is too deep for the device evaluator");
        if (hdr.error >= 0x10000u) {
            const unsigned id = hdr.error - 0x10000u;
            throw std::runtime_error("Assertion error. The expression of the tool \"" +
                                     (id < tools.size() ? tools[id]->name() : std::string("?")) +
                                     "\" is false");
        }
        throw std::runtime_error("A value computed in " + where + " overflows its variable type");
    }
    return recorded;
}

// ------------------------------------------------- the tools' side of it --
bool Kernel::isSweep() const
{
    for (auto v : _vars)
        if (v->name() == "ihoc" || v->name() == "mpi_ihoc")
            return true;
    return false;
}

size_t Kernel::globalSize() const
{
    // global size: n="" -> longest array argument (Kernel.cpp:558-594)
    size_t N = 0;
    if (_n.empty()) {
        for (auto v : _vars)
            if (v->isArray() && v->length() > N)
                N = v->length();
    } else {
        uint64_t n = 0;
        _C->variables()->solve("unsigned long", _n, &n);
        N = (size_t)n;
    }
    return N;
}

bool Kernel::recordable(DeviceLoop& L, std::string& why) const
{
    if (_leader) {
        if (!L.contains(_leader)) {
            why = "its fused group starts outside the loop";
            return false;
        }
        return true;
    }
    if (_fused_id >= 0) {
        for (auto k : _group) {
            if (!L.contains(k)) {
                why = "its fused group leaves the loop";
                return false;
            }
            for (auto v : k->_vars)
                if (!v->isArray() && L.varying(v)) {
                    why = "a fused sweep reads the loop scalar \"" + v->name() + "\" by value";
                    return false;
                }
        }
        return true;
    }
    const uint64_t mask = aqc_kernel_dev_scalars(_kid);
    for (size_t k = 0; k < _vars.size(); k++)
        if (!_vars[k]->isArray() && L.varying(_vars[k])) {
            if (k >= 64 || !((mask >> k) & 1) || _vars[k]->typesize() != 4) {
                why = "the kernel takes the loop scalar \"" + _vars[k]->name() + "\" by value";
                return false;
            }
        }
    if (!_n.empty() && L.readsVarying(_n)) {
        why = "its global size depends on a loop scalar";
        return false;
    }
    return true;
}

void Kernel::record(DeviceLoop& L)
{
    if (_leader)
        return;
    // (scalar programs queued so far run in front of the next kernel that READS the table: the
    // order among programs is what matters to them, and nothing else touches the table)
    if (_fused_id >= 0) {
        _execute();
        return;
    }
    const size_t N = globalSize();
    std::vector<void*> args(_vars.size());
    std::vector<const void*> dev(_vars.size(), nullptr);
    bool any = false;
    for (size_t k = 0; k < _vars.size(); k++) {
        args[k] = _vars[k]->isArray() ? _vars[k]->dptr() : _vars[k]->get();
        if (!_vars[k]->isArray() && L.varying(_vars[k])) {
            dev[k] = L.deviceAddress(_vars[k]);
            any = true;
        }
    }
    if (any)
        L.flush();
    check(aqc_launch_ex(_C->ctx(), _kid, N, args.data(), (int)args.size(), any ? dev.data() : nullptr));
}

void Copy::record(DeviceLoop& L)
{
    (void)L;
    _execute();
}

bool Set::recordable(DeviceLoop& L, std::string& why) const
{
    bool reads = false;
    try {
        reads = L.readsVarying(_value);
    } catch (std::exception&) {
        reads = false; // a literal / macro (VEC_ZERO ...): no variable involved
    }
    if (reads) {
        why = "its value depends on a loop scalar";
        return false;
    }
    return true;
}

void Set::record(DeviceLoop& L)
{
    (void)L;
    _execute();
}

bool SetScalar::recordable(DeviceLoop& L, std::string& why) const
{
    if (!L.varying(_var)) {
        why = "its variable is not in the loop's table";
        return false;
    }
    if (split_formulae(_expr).size() < _var->ncomp()) {
        why = "too few components";
        return false;
    }
    return L.compilable(_expr, why);
}

void SetScalar::record(DeviceLoop& L)
{
    // Variables::solve: every component is evaluated before the variable changes
    const auto parts = split_formulae(_expr);
    const unsigned n = _var->ncomp();
    std::vector<aqs_op> ops;
    for (unsigned c = 0; c < n; c++)
        L.compile(parts[c], ops);
    const int off = L.offset(_var);
    for (unsigned c = n; c-- > 0;)
        ops.push_back(mk(AQS_STORE, off + 4 * (int)c, _var->kind()));
    L.emit(ops);
}

bool Assert::recordable(DeviceLoop& L, std::string& why) const { return L.compilable(_expr, why); }

void Assert::record(DeviceLoop& L)
{
    std::vector<aqs_op> ops;
    L.compile(_expr, ops);
    ops.push_back(mk(AQS_ASSERT, id_in_pipeline()));
    L.emit(ops);
}

bool Reduction::recordable(DeviceLoop& L, std::string& why) const
{
    if (!L.varying(_out)) {
        why = "its output is not in the loop's table";
        return false;
    }
    // the table slots of its own are asked for here, while the table is still being laid out
    L.scratch(this, 0);
    const int ident = L.scratch(this, 1);
    L.setInitial(ident, _identity.data(), std::min<size_t>(16, _identity.size()));
    return true;
}

void Reduction::record(DeviceLoop& L)
{
    const int raw = L.scratch(this, 0), ident = L.scratch(this, 1);
    check(aqc_reduce(_C->ctx(), _op, _atype, _in->dptr(), _in->length(), L.scratchDevice(raw), nullptr));
    // fold the user's null value in, in the array's own arithmetic (Reduction::_execute)
    const int kc = _in->kind() == 'f' ? 0 : (_in->kind() == 'u' ? 1 : 2);
    L.emit(mk(AQS_FOLD, L.offset(_out), raw, ident, (double)(_op + 4 * kc + 16 * (int)_in->ncomp())));
}

bool Report::recordable(DeviceLoop&, std::string& why) const
{
    if (_kind != "screen" && _kind != "file") {
        why = "report_" + _kind + " reads the device";
        return false;
    }
    for (auto v : _vars)
        if (v->isArray()) {
            why = "it prints the array \"" + v->name() + "\"";
            return false;
        }
    return true;
}

void Report::record(DeviceLoop& L) { L.emit(mk(AQS_SNAP, id_in_pipeline())); }

// ------------------------------------------------------------------- While --
While::~While() { delete _dev; }

void While::planDeviceLoop()
{
    if (const char* e = getenv("AQUA_DEVICE_LOOPS"))
        if (!strcmp(e, "0")) {
            _why = "AQUA_DEVICE_LOOPS=0";
            return;
        }
    if (getenv("AQUA_PROFILE_SYNC")) {
        _why = "AQUA_PROFILE_SYNC times every tool on the host";
        return;
    }
    if (once()) {
        _why = "once=\"true\"";
        return;
    }
    const size_t first = (size_t)id_in_pipeline() + 1;
    const size_t last = _ending_tool ? (size_t)_ending_tool->id_in_pipeline() - 1 : _C->tools().size() - 1;
    _dev = new DeviceLoop(_C, this, _expr, first, last);
    if (!_dev->plan(_why)) {
        delete _dev;
        _dev = nullptr;
        log(L_INFO, "The loop \"" + name() + "\" runs on the host: " + _why + "\n");
        return;
    }
    log(L_INFO, "The loop \"" + name() + "\" runs on the device (" + std::to_string(last - first) +
                    " tools: first pass tool by tool without read-backs, the rest as a CUDA graph "
                    "while-node)\n");
}

void While::_execute()
{
    const bool again = _reentry;
    _reentry = false;
    Conditional::_execute();
    if (again || !_result || !_dev || !_dev->usable())
        return;
    // entering the loop: all of it on the device
    if (_dev->run()) {
        _result = false; // the loop is over: on to the tool behind its `end`
        return;
    }
    // the body could not be recorded: its first pass did run (on device-resident scalars, now back
    // on the host) and the loop goes on tool by tool from its condition
    Conditional::_execute();
}

Tool* End::next_tool()
{
    Tool* opener = Tool::next_tool();
    if (While* w = dynamic_cast<While*>(opener))
        w->reentry();
    return opener;
}

void CalcServer::planDeviceLoops()
{
    _device_loops = 0;
    for (auto& t : _tools)
        if (While* w = dynamic_cast<While*>(t.get())) {
            w->planDeviceLoop();
            if (w->deviceLoop())
                _device_loops++;
        }
}

} // namespace CalcServer
} // namespace Aqua
