// capi.cpp -- C entry points of libaquahost.so (include/aquahost.h)
#include <cstdlib>
#include <cstring>
#include <string>

#include "aquahost.h"
#include "calcserver.hpp"
#include "devloop.hpp"

using namespace Aqua;

struct aqh_sim {
    InputOutput::ProblemSetup sim_data;
    std::unique_ptr<CalcServer::CalcServer> C;
    std::unique_ptr<CalcServer::TimeManager> T;
    std::vector<std::string> tool_types, tool_names;
};

static thread_local std::string g_err;

// AQUA_SEGV_BACKTRACE=1: a fatal signal inside the host or a plugin prints the native frames first
// (the driving process usually only knows its own, interpreted, ones)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void aqh_fatal_signal(int sig)
{
    const char msg[] = "libaquahost: fatal signal, native backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    void* frames[96];
    const int n = backtrace(frames, 96);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
namespace {
struct SegvBacktrace {
    SegvBacktrace()
    {
        const char* e = getenv("AQUA_SEGV_BACKTRACE");
        if (!e || atoi(e) == 0)
            return;
        // an alternate stack (a stack overflow leaves no room for a handler on the faulting one) and
        // one backtrace() up front (its first call loads libgcc, which a signal handler must not do)
        static char altstack[1 << 16];
        stack_t ss;
        ss.ss_sp = altstack;
        ss.ss_size = sizeof(altstack);
        ss.ss_flags = 0;
        sigaltstack(&ss, nullptr);
        void* warm[4];
        backtrace(warm, 4);
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_handler = aqh_fatal_signal;
        sa.sa_flags = SA_ONSTACK | SA_NODEFER;
        sigemptyset(&sa.sa_mask);
        sigaction(SIGSEGV, &sa, nullptr);
        sigaction(SIGBUS, &sa, nullptr);
        sigaction(SIGABRT, &sa, nullptr);
    }
} g_segv_backtrace;
} // namespace

#define AQH_TRY try {
#define AQH_CATCH                                                              \
    }                                                                          \
    catch (std::exception & e)                                                 \
    {                                                                          \
        g_err = e.what();                                                      \
        return -1;                                                             \
    }                                                                          \
    catch (...)                                                                \
    {                                                                          \
        g_err = "unknown error";                                               \
        return -1;                                                             \
    }

extern "C" const char* aqh_last_error(void) { return g_err.c_str(); }
extern "C" void aqh_set_log_level(int level) { logLevel() = level; }

static std::unique_ptr<aqh_sim> parse_only(const char* xml_path, int dims, const char* root_path)
{
    if (dims != 2 && dims != 3)
        throw std::runtime_error("dims must be 2 or 3");
    auto sim = std::make_unique<aqh_sim>();
    sim->sim_data.dims = dims;
    if (root_path && *root_path)
        sim->sim_data.settings.base_path = root_path;
    else if (getenv("AQUAGPUSPH_ROOT"))
        sim->sim_data.settings.base_path = getenv("AQUAGPUSPH_ROOT");
    InputOutput::State state;
    state.load(xml_path, sim->sim_data);
    for (auto& t : sim->sim_data.tools) {
        sim->tool_types.push_back(t->get("type"));
        sim->tool_names.push_back(t->get("name"));
    }
    for (auto& r : sim->sim_data.reports) {
        sim->tool_types.push_back("report_" + r->get("type"));
        sim->tool_names.push_back(r->get("name"));
    }
    return sim;
}

extern "C" int aqh_parse(const char* xml_path, int dims, const char* root_path, aqh_sim** out)
{
    if (!out || !xml_path) {
        g_err = "aqh_parse: NULL argument";
        return -1;
    }
    *out = nullptr;
    AQH_TRY
    *out = parse_only(xml_path, dims, root_path).release();
    return 0;
    AQH_CATCH
}

extern "C" int aqh_load(const char* xml_path, int dims, int device, const char* root_path,
                        int mpi_rank, int mpi_size, aqh_sim** out)
{
    if (!out || !xml_path) {
        g_err = "aqh_load: NULL argument";
        return -1;
    }
    *out = nullptr;
    AQH_TRY
    auto sim = parse_only(xml_path, dims, root_path);
    sim->C = std::make_unique<CalcServer::CalcServer>(sim->sim_data, device, mpi_rank, mpi_size);
    sim->C->loadParticles();
    sim->C->setup();
    sim->T = std::make_unique<CalcServer::TimeManager>(sim->C.get(), sim->sim_data);
    *out = sim.release();
    return 0;
    AQH_CATCH
}

extern "C" int aqh_comm_unique_id(void* id_out)
{
    if (aqc_comm_unique_id(id_out)) {
        g_err = "cannot create a NCCL unique id (libnccl.so.2 missing?)";
        return -1;
    }
    return 0;
}

extern "C" int aqh_comm_init(aqh_sim* sim, const void* unique_id)
{
    AQH_TRY
    sim->C->commInit(unique_id);
    return 0;
    AQH_CATCH
}

extern "C" void aqh_destroy(aqh_sim* sim) { delete sim; }

extern "C" int aqh_write_resolved(aqh_sim* sim, const char* path)
{
    AQH_TRY
    InputOutput::State().write(path, sim->sim_data);
    return 0;
    AQH_CATCH
}

extern "C" int aqh_n_tools(aqh_sim* sim) { return sim ? (int)sim->tool_names.size() : 0; }
extern "C" const char* aqh_tool_name(aqh_sim* sim, int i)
{
    return (i >= 0 && (size_t)i < sim->tool_names.size()) ? sim->tool_names[i].c_str() : nullptr;
}
extern "C" const char* aqh_tool_type(aqh_sim* sim, int i)
{
    return (i >= 0 && (size_t)i < sim->tool_types.size()) ? sim->tool_types[i].c_str() : nullptr;
}
extern "C" double aqh_tool_elapsed_ms(aqh_sim* sim, int i)
{
    auto* t = sim->C ? sim->C->tool(i) : nullptr;
    return t ? t->elapsed_ms() : 0.0;
}
extern "C" unsigned aqh_tool_used_times(aqh_sim* sim, int i)
{
    auto* t = sim->C ? sim->C->tool(i) : nullptr;
    return t ? t->used_times() : 0;
}

extern "C" int aqh_step(aqh_sim* sim, int n)
{
    AQH_TRY
    for (int k = 0; k < n; k++)
        sim->C->step();
    return 0;
    AQH_CATCH
}

extern "C" int aqh_run(aqh_sim* sim)
{
    AQH_TRY
    // main.cpp:162-181: update until an output frame is due, save it, ...; wait for the writers
    while (!sim->T->mustStop()) {
        sim->C->update(*sim->T);
        sim->C->save(sim->T->time());
    }
    if (aqc_sync(sim->C->ctx()))
        throw std::runtime_error(aqc_last_error(sim->C->ctx()));
    sim->C->waitForSavers();
    return 0;
    AQH_CATCH
}

extern "C" int aqh_save(aqh_sim* sim)
{
    AQH_TRY
    sim->C->save(sim->T->time());
    return 0;
    AQH_CATCH
}

extern "C" int aqh_wait_savers(aqh_sim* sim)
{
    AQH_TRY
    sim->C->waitForSavers();
    return 0;
    AQH_CATCH
}

extern "C" const char* aqh_checkpoint_file(aqh_sim* sim)
{
    return (sim && sim->C) ? sim->C->checkpoint_file().c_str() : "";
}

extern "C" int aqh_n_savers(aqh_sim* sim) { return (sim && sim->C) ? (int)sim->C->savers().size() : 0; }
extern "C" const char* aqh_saver_file(aqh_sim* sim, int i)
{
    if (!sim || !sim->C || i < 0 || (size_t)i >= sim->C->savers().size())
        return "";
    return sim->C->savers()[i]->file().c_str();
}

extern "C" int aqh_sync(aqh_sim* sim)
{
    AQH_TRY
    if (aqc_sync(sim->C->ctx()))
        throw std::runtime_error(aqc_last_error(sim->C->ctx()));
    return 0;
    AQH_CATCH
}

extern "C" uint64_t aqh_launch_count(aqh_sim* sim) { return aqc_launch_count(sim->C->ctx()); }
extern "C" void* aqh_cuda_ctx(aqh_sim* sim) { return sim->C->ctx(); }
extern "C" unsigned aqh_fused_groups(aqh_sim* sim) { return sim->C ? sim->C->fused_groups() : 0; }

static void declareScalars(InputOutput::Variables& vars, const char* decls, const char* who)
{
    std::string d(decls ? decls : "");
    size_t pos = 0;
    while (pos < d.size()) {
        size_t end = d.find(';', pos);
        if (end == std::string::npos)
            end = d.size();
        const std::string item = trimCopy(d.substr(pos, end - pos));
        pos = end + 1;
        if (item.empty())
            continue;
        const size_t eq = item.find('=');
        if (eq == std::string::npos)
            throw std::runtime_error(std::string(who) + ": \"" + item + "\" is not \"type name=value\"");
        const std::string lhs = trimCopy(item.substr(0, eq));
        const size_t sp = lhs.find_last_of(" \t");
        if (sp == std::string::npos)
            throw std::runtime_error(std::string(who) + ": \"" + item + "\" is not \"type name=value\"");
        vars.registerVariable(trimCopy(lhs.substr(sp + 1)), trimCopy(lhs.substr(0, sp)), "",
                              trimCopy(item.substr(eq + 1)));
    }
}

extern "C" int aqh_eval(int dims, const char* decls, const char* type, const char* expr,
                        void* out, size_t bytes)
{
    AQH_TRY
    if (!type || !expr || !out)
        throw std::runtime_error("aqh_eval: NULL argument");
    InputOutput::Variables vars(dims, nullptr);
    declareScalars(vars, decls, "aqh_eval");
    vars.exprVariables(expr); // unknown names are an error (Variable.cpp:1230-1237)
    const size_t ts = vars.typeToBytes(type);
    if (!ts || ts > bytes)
        throw std::runtime_error(std::string("aqh_eval: bad type or buffer for \"") + type + "\"");
    vars.solve(type, expr, out);
    return 0;
    AQH_CATCH
}

// The same value through the device evaluator's code path, on the host: every declared 32-bit
// scalar lives in a table (as the loop scalars of a recorded `while` do), `expr` is compiled by
// SvmCompiler and run by aqs_run (include/aquasvm.h, the source the one-thread kernel of
// csrc/devloop.cu is built from)
namespace {
struct SvmTable {
    InputOutput::Variables* vars;
    std::map<const InputOutput::Variable*, int> slots;
    static bool resolve(void* user, const std::string& id, CalcServer::SvmCompiler::Slot& out)
    {
        SvmTable* T = (SvmTable*)user;
        InputOutput::Variable* v = T->vars->get(id);
        unsigned comp = 0;
        if (!v) {
            static const char* suf[] = { "_x", "_y", "_z", "_w" };
            for (unsigned c = 0; c < 4 && !v; c++)
                if (endswith(id, suf[c])) {
                    InputOutput::Variable* b = T->vars->get(id.substr(0, id.size() - 2));
                    if (b && b->ncomp() > c) {
                        v = b;
                        comp = c;
                    }
                }
            if (!v)
                return false;
        } else if (v->ncomp() != 1)
            return false;
        auto it = T->slots.find(v);
        if (it == T->slots.end())
            return false;
        out.offset = it->second + 4 * (int)comp;
        out.kind = v->kind();
        return true;
    }
};
} // namespace

extern "C" int aqh_eval_svm(int dims, const char* decls, const char* type, const char* expr,
                            void* out, size_t bytes)
{
    AQH_TRY
    if (!type || !expr || !out)
        throw std::runtime_error("aqh_eval_svm: NULL argument");
    InputOutput::Variables vars(dims, nullptr);
    declareScalars(vars, decls, "aqh_eval_svm");
    vars.exprVariables(expr);
    const size_t ts = vars.typeToBytes(type);
    if (!ts || ts > bytes || ts > 64)
        throw std::runtime_error(std::string("aqh_eval_svm: bad type or buffer for \"") + type + "\"");
    // the result type as the host describes it
    vars.registerVariable("__aqh_result__", type, "", "");
    InputOutput::Variable* res = vars.get("__aqh_result__");
    if (res->kind() != 'f' && res->kind() != 'u' && res->kind() != 'i')
        throw std::runtime_error("aqh_eval_svm: only 32-bit component types live on the device");
    SvmTable T{ &vars, {} };
    int table_bytes = 0;
    for (auto& v : vars.all())
        if (!v->isArray() && (v->kind() == 'f' || v->kind() == 'u' || v->kind() == 'i') &&
            v->typesize() <= 64) {
            T.slots[v.get()] = table_bytes;
            table_bytes += 64;
        }
    std::vector<char> table(table_bytes, 0);
    for (auto& kv : T.slots)
        memcpy(table.data() + kv.second, kv.first->get(), kv.first->typesize());
    CalcServer::SvmCompiler comp(&vars.tokenizer(), &SvmTable::resolve, &T);
    const auto parts = split_formulae(expr);
    const unsigned n = res->ncomp();
    if (parts.size() < n)
        throw std::runtime_error("Invalid number of fields in \"" + std::string(expr) + "\"");
    std::vector<aqs_op> prog;
    for (unsigned c = 0; c < n; c++)
        comp.compile(parts[c], prog);
    const int off = T.slots[res];
    for (unsigned c = n; c-- > 0;) {
        aqs_op o;
        o.code = AQS_STORE;
        o.a = off + 4 * (int)c;
        o.b = res->kind();
        o.c = 0;
        o.imm = 0.0;
        prog.push_back(o);
    }
    aqs_header hdr;
    memset(&hdr, 0, sizeof(hdr));
    aqs_run(prog.data(), (int)prog.size(), table.data(), table_bytes, &hdr, nullptr, 0, 1u);
    if (hdr.error)
        throw std::out_of_range("aqh_eval_svm: error " + std::to_string(hdr.error) +
                                " (the value overflows the variable type)");
    memcpy(out, table.data() + off, ts);
    return 0;
    AQH_CATCH
}

extern "C" unsigned aqh_device_loops(aqh_sim* sim) { return sim && sim->C ? sim->C->device_loops() : 0; }

extern "C" int aqh_device_loop_stats(aqh_sim* sim, uint64_t* runs, uint64_t* iterations)
{
    AQH_TRY
    uint64_t r = 0, it = 0;
    for (auto& t : sim->C->tools())
        if (auto* w = dynamic_cast<CalcServer::While*>(t.get()))
            if (w->deviceLoop()) {
                r += w->deviceLoop()->runs();
                it += w->deviceLoop()->iterations();
            }
    if (runs)
        *runs = r;
    if (iterations)
        *iterations = it;
    return 0;
    AQH_CATCH
}

extern "C" int aqh_device_loop_timing(aqh_sim* sim, int* body_nodes, double* record_ms, double* instantiate_ms)
{
    AQH_TRY
    int n = 0;
    double r = 0.0, i = 0.0;
    for (auto& t : sim->C->tools())
        if (auto* w = dynamic_cast<CalcServer::While*>(t.get()))
            if (w->deviceLoop()) {
                int n1;
                double r1, i1;
                w->deviceLoop()->timing(n1, r1, i1);
                n += n1;
                r += r1;
                i += i1;
            }
    if (body_nodes)
        *body_nodes = n;
    if (record_ms)
        *record_ms = r;
    if (instantiate_ms)
        *instantiate_ms = i;
    return 0;
    AQH_CATCH
}

extern "C" int aqh_lane_schedule(int n, const int* r_off, const int* r_var, const unsigned* r_rows, const int* w_off,
                                 const int* w_var, const unsigned* w_rows, const double* cost,
                                 const unsigned char* flags, double gain, int* lane_out, int* wait_out,
                                 unsigned char* marked_out, int* last_lane1_out)
{
    AQH_TRY
    if (n < 0 || !r_off || !w_off || !cost || !flags || !lane_out || !wait_out || !marked_out)
        throw std::runtime_error("aqh_lane_schedule: NULL argument");
    std::vector<CalcServer::LaneDep> deps(n);
    for (int k = 0; k < n; k++) {
        for (int e = r_off[k]; e < r_off[k + 1]; e++)
            deps[k].r.emplace_back((const void*)(intptr_t)(r_var[e] + 1), r_rows[e]);
        for (int e = w_off[k]; e < w_off[k + 1]; e++)
            deps[k].w.emplace_back((const void*)(intptr_t)(w_var[e] + 1), w_rows[e]);
        deps[k].cost = cost[k];
        deps[k].forced0 = (flags[k] & 1) != 0;
        deps[k].barrier = (flags[k] & 2) != 0;
        deps[k].launches = (flags[k] & 4) == 0;
    }
    std::vector<int> lane;
    std::vector<std::vector<int>> waits;
    std::vector<char> marked;
    int last1 = -1;
    CalcServer::scheduleLanes(deps, gain, lane, waits, marked, last1);
    for (int k = 0; k < n; k++) {
        lane_out[k] = lane[k];
        wait_out[k] = waits[k].empty() ? -1 : waits[k].back();
        marked_out[k] = (unsigned char)marked[k];
    }
    if (last_lane1_out)
        *last_lane1_out = last1;
    return 0;
    AQH_CATCH
}

extern "C" unsigned aqh_device_loop_branch_tools(aqh_sim* sim)
{
    unsigned n = 0;
    if (sim && sim->C)
        for (auto& t : sim->C->tools())
            if (auto* w = dynamic_cast<CalcServer::While*>(t.get()))
                if (w->deviceLoop())
                    n += w->deviceLoop()->branchTools();
    return n;
}

extern "C" const char* aqh_loop_host_reason(aqh_sim* sim, int i)
{
    if (!sim || !sim->C || i < 0 || i >= (int)sim->C->tools().size())
        return nullptr;
    auto* w = dynamic_cast<CalcServer::While*>(sim->C->tools()[i].get());
    if (!w)
        return nullptr;
    static thread_local std::string why;
    why = w->deviceLoop() ? "" : w->hostReason();
    return why.c_str();
}

extern "C" void aqh_set_script_runner(aqh_script_fn fn, void* user)
{
    Aqua::CalcServer::setScriptRunner(fn, user);
}

extern "C" const char* aqh_variable_type(aqh_sim* sim, const char* name)
{
    if (!sim || !sim->C || !name)
        return nullptr;
    auto* v = sim->C->variables()->get(name);
    if (!v)
        return nullptr;
    static thread_local std::string type;
    type = v->type();
    return type.c_str();
}

extern "C" int aqh_scalar_get(aqh_sim* sim, const char* name, void* out, size_t bytes)
{
    AQH_TRY
    auto* v = sim->C->variables()->get(name);
    if (!v || v->isArray())
        throw std::runtime_error(std::string("No such scalar variable \"") + name + "\"");
    if (bytes < v->typesize())
        throw std::runtime_error(std::string("Buffer too small for \"") + name + "\"");
    memcpy(out, v->get(), v->typesize());
    return 0;
    AQH_CATCH
}

extern "C" int aqh_scalar_set(aqh_sim* sim, const char* name, const char* expression)
{
    AQH_TRY
    auto* v = sim->C->variables()->get(name);
    if (!v || v->isArray())
        throw std::runtime_error(std::string("No such scalar variable \"") + name + "\"");
    sim->C->variables()->solve(v->type(), expression, v->get(), name);
    return 0;
    AQH_CATCH
}

extern "C" int aqh_array_info(aqh_sim* sim, const char* name, size_t* length, size_t* elem_bytes)
{
    AQH_TRY
    auto* v = sim->C->variables()->get(name);
    if (!v || !v->isArray())
        throw std::runtime_error(std::string("No such array variable \"") + name + "\"");
    if (length)
        *length = v->length();
    if (elem_bytes)
        *elem_bytes = v->typesize();
    return 0;
    AQH_CATCH
}

extern "C" int aqh_array_download(aqh_sim* sim, const char* name, void* host_out, int unsorted)
{
    AQH_TRY
    if (unsorted)
        sim->C->getUnsortedMem(name, host_out);
    else
        sim->C->download(name, host_out);
    return 0;
    AQH_CATCH
}

extern "C" int aqh_array_upload(aqh_sim* sim, const char* name, const void* host_in)
{
    AQH_TRY
    sim->C->upload(name, host_in);
    return 0;
    AQH_CATCH
}

extern "C" void* aqh_array_devptr(aqh_sim* sim, const char* name)
{
    auto* v = sim->C->variables()->get(name);
    return (v && v->isArray()) ? v->dptr() : nullptr;
}
