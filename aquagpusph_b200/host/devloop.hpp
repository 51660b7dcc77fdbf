// devloop.hpp -- a `while` loop of the pipeline run on the device.
//
// The reference's while / set_scalar / reduction / report tools meet on the host
// (Conditional.cpp:85-96, SetScalar.cpp:146-195, Reduction.cpp:205-258): every sub-iteration of the
// midpoint scheme waits for its reductions before the host can evaluate the relaxation and stop
// expressions and enqueue the next pass.  When every tool between a `while` and its `end` can be
// RECORDED -- it only enqueues device work, or it is scalar arithmetic over the tokenizer's
// variables -- the loop runs as one CUDA graph with a WHILE node (csrc/devloop.cu):
//   * the scalar variables the body writes live in a table in device memory,
//   * set_scalar / assert / the loop condition are compiled by SvmCompiler (the grammar of
//     host/tokenizer.hpp) into the stack programs of include/aquasvm.h,
//   * reductions leave their result in the table and are folded with their null value there,
//   * kernels that read a scalar the loop writes take it from the table (aqc_launch_ex),
//   * report tools snapshot the table; the host prints the snapshots when the loop is over.
// The first pass of every loop still runs tool by tool (it is the pass that builds the neighbour
// lists and sizes every scratch buffer), but already on the device table: nothing is read back, so
// the device is busy with it while the host records and instantiates the body for the passes that
// follow; the graph starts from the condition the first pass left.  One synchronisation per loop.
// A body that cannot be recorded -- link-list, mpi-sync, python, nested conditionals, 64-bit
// scalars -- stays on the host path, unchanged; a recording that fails at run time hands the
// scalars back after the first pass and the loop goes on tool by tool.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "aquacuda.h"
#include "aquasvm.h"
#include "tokenizer.hpp"
#include "variables.hpp"

namespace Aqua {
namespace CalcServer {

class CalcServer;
class Tool;

/// Expression -> stack program.  Mirrors Tokenizer::P production by production (same precedence,
/// same associativity, both branches of ?: evaluated); identifiers resolve through `slot` to a
/// table entry, else to the tokenizer's current value as an immediate.
class SvmCompiler {
  public:
    struct Slot {
        int offset;
        char kind; // 'f', 'u', 'i'
    };
    /// name (a scalar variable, or name_x.. of a vector one) -> table slot; false: not in the table
    typedef bool (*Resolver)(void* user, const std::string& id, Slot& out);
    SvmCompiler(const Tokenizer* tok, Resolver r, void* user) : _tok(tok), _resolve(r), _user(user) {}
    /// appends the ops that leave the value of `expr` on the stack; throws on a syntax error, an
    /// unknown name or function
    void compile(const std::string& expr, std::vector<aqs_op>& out) const;

  private:
    const Tokenizer* _tok;
    Resolver _resolve;
    void* _user;
};

/// What one tool of a loop body reads and writes, for the two-lane schedule: opaque array keys with the
/// particle classes (AQC_ROWS_*) of the rows involved
struct LaneDep {
    typedef std::vector<std::pair<const void*, unsigned>> Access;
    Access r, w;
    bool barrier = false;  ///< unknown dependencies: conflicts with everything
    bool forced0 = false;  ///< must run on lane 0 (reductions, kernels reading the scalar table)
    bool launches = true;  ///< false: nothing is enqueued at this position (scalar tools, fused followers)
    double cost = 1.0;
};
/// lane_of[k] in {0, 1}; waits[k]: tools of the other lane whose event tool k waits for; marked[k]: an event is
/// recorded behind tool k; last_lane1: the last tool on lane 1 (-1: none), joined at the end of the pass
void scheduleLanes(const std::vector<LaneDep>& deps, double gain, std::vector<int>& lane_of,
                   std::vector<std::vector<int>>& waits, std::vector<char>& marked, int& last_lane1);

class DeviceLoop {
  public:
    /// body: the tools between the opening `while` (index `first` - 1) and its `end` (index `last`)
    DeviceLoop(CalcServer* C, Tool* opener, const std::string& condition, size_t first, size_t last);
    ~DeviceLoop();
    /// Classify the body and lay the table out.  false: the loop stays on the host (`why` says)
    bool plan(std::string& why);
    bool usable() const { return _usable; }
    /// Run the loop (the host found its condition true): true when it is over (variables updated,
    /// reports printed); false when only the first pass ran and the host has to carry on tool by
    /// tool from the condition
    bool run();
    uint64_t runs() const { return _runs; }
    uint64_t iterations() const { return _iterations; }
    /// kernel / copy / memset nodes of the recorded body, host milliseconds of the last recording
    /// and of the last instantiation
    void timing(int& body_nodes, double& record_ms, double& instantiate_ms) const
    {
        body_nodes = 0;
        record_ms = instantiate_ms = 0.0;
        if (_loop)
            aqc_loop_stats(_loop, &body_nodes, &record_ms, &instantiate_ms);
    }

    // ---- what the tools see while they are asked (recordable) or recorded (record)
    bool contains(const Tool* t) const;
    bool varying(const InputOutput::Variable* v) const { return _slots.count(v) > 0; }
    /// offset of a variable in the table (-1: not part of it)
    int offset(const InputOutput::Variable* v) const;
    const void* deviceAddress(const InputOutput::Variable* v) const;
    /// true when `expr` only uses names a program can read and compiles
    bool compilable(const std::string& expr, std::string& why) const;
    bool readsVarying(const std::string& expr) const;
    void compile(const std::string& expr, std::vector<aqs_op>& out) const;
    /// 16 bytes of the table of the tool's own (reductions: raw result, null value)
    int scratch(const Tool* t, int which);
    void* scratchDevice(int offset) const;
    void setInitial(int offset, const void* data, size_t bytes);
    void emit(const aqs_op& op) { _pending.push_back(op); }
    void emit(const std::vector<aqs_op>& ops) { _pending.insert(_pending.end(), ops.begin(), ops.end()); }
    /// the scalar programs queued so far become one kernel in front of the next device tool
    void flush();
    aqc_ctx* ctx() const;

  private:
    static bool resolve(void* user, const std::string& id, SvmCompiler::Slot& out);
    void planLanes();
    void lane(int l);
    void pass();
    void record();
    CalcServer* _C;
    Tool* _opener;
    std::string _condition;
    size_t _first, _last;
    bool _usable = false, _planned = false;
    std::map<const InputOutput::Variable*, int> _slots;
    std::vector<InputOutput::Variable*> _order; // table variables in layout order
    std::map<std::pair<const Tool*, int>, int> _scratch;
    int _table_bytes = 0;
    std::vector<char> _initial; // scratch constants (null values)
    std::vector<aqs_op> _pending;
    aqc_loop* _loop = nullptr;
    int _hist_rows = 0;
    unsigned _failures = 0;
    uint64_t _runs = 0, _iterations = 0;
    uint32_t _max_iters = 65536;
    // two lanes (aqc_lane_*): tools whose arrays do not meet run on a second stream, ordered against
    // the first by events behind the tools they depend on -- what the reference's queue pool and
    // per-variable events do (Tool.cpp:405-444), decided once, at plan time
    bool _two_lanes = false;
    int _cur_lane = 0;
    std::vector<int> _lane_of;            // per body tool
    std::vector<std::vector<int>> _waits; // per body tool: body tools (other lane) whose event it waits for
    std::vector<char> _marked;            // per body tool: an event is recorded behind it
    std::vector<void*> _events;           // per body tool (created on first use)
    void* _ev_fork = nullptr;
    int _last_lane1 = -1;
  public:
    /// tools of the body that run on the branch lane (0: single lane)
    unsigned branchTools() const;
};

} // namespace CalcServer
} // namespace Aqua
