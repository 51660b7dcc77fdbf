// calcserver.hpp -- the scheduler and the tool types of the C++ host.
//
// Counterpart of aquagpusph/CalcServer/{CalcServer,Tool,Kernel,LinkList,
// RadixSort,UnSort,Reduction,Set,SetScalar,Copy,Conditional,Assert}.* and
// Reports/{Screen,TabFile,Dump}: same XML tool types and attributes, same
// per-step pipeline semantics (CalcServer::update, CalcServer.cpp:592-621), but
// every device operation goes through the C-ABI of libaquacuda.so
// (include/aquacuda.h) on ONE in-order CUDA stream instead of a pool of OpenCL
// queues stitched together with events.  Stream order subsumes the
// read/write-event graph of Tool.cpp:405-444; the host blocks only where a
// scalar computed on the device is needed by host logic (reductions, the
// link-list grid), which is where the reference blocks too.
#pragma once
#include <thread>
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "aquacuda.h"
#include "problem.hpp"
#include "variables.hpp"

namespace Aqua {
namespace CalcServer {

class CalcServer;
class DeviceLoop;

/// Base class of every tool (Tool.hpp:173-567)
class Tool {
  public:
    Tool(CalcServer* C, const std::string& name, bool once = false)
      : _C(C), _name(name), _once(once) {}
    /// The reference's constructor (Tool.hpp:180: name, once): what a type="installable" plugin's
    /// create_object(name, once) calls; the host attaches the server right after
    Tool(const std::string& name, bool once = false) : _C(nullptr), _name(name), _once(once) {}
    void attach(CalcServer* C) { _C = C; }
    virtual ~Tool() {}
    const std::string& name() const { return _name; }
    virtual void setup() {}
    /// Run the tool (honouring once="true") and account its host time
    void execute();
    /// Tool to run next; default: the following one in the pipeline
    virtual Tool* next_tool() { return _next; }
    void next_tool(Tool* t) { _next = t; }
    /// +1 for tools opening a scope (if/while), -1 for end (Tool.hpp:351-357)
    virtual int scope_modifier() const { return 0; }
    void id_in_pipeline(int i) { _id = i; }
    int id_in_pipeline() const { return _id; }
    unsigned used_times() const { return _n_iters; }
    double elapsed_ms() const { return _elapsed_ms; }
    /// Variables the tool reads / writes (Tool::setDependencies, Tool.hpp:520-567).
    /// false: unknown -- the tool is a barrier for the sweep-fusion planner.
    virtual bool dependencies(std::vector<InputOutput::Variable*>& in,
                              std::vector<InputOutput::Variable*>& out) const
    {
        (void)in; (void)out;
        return false;
    }
    // ---- device-side loops (host/devloop.hpp).  recordable: one pass of the tool inside the loop
    // L only enqueues device work or is scalar arithmetic a program can do (no read-back, no host
    // decision) -- asked once, while the loop's table is laid out: a tool may reserve slots of its
    // own there (DeviceLoop::scratch); scalarOutputs: the scalar variables it writes (they move into the loop's device
    // table); record: enqueue / emit that pass while the loop's body is being recorded.
    virtual bool recordable(DeviceLoop& L, std::string& why) const
    {
        (void)L;
        why = "its tool type runs on the host";
        return false;
    }
    virtual void scalarOutputs(std::vector<InputOutput::Variable*>& out) const { (void)out; }
    virtual void record(DeviceLoop& L) { (void)L; }
    bool once() const { return _once; }
    /// passes that ran inside a device-side loop
    void account(unsigned n) { _n_iters += n; }

  protected:
    virtual void _execute() {}
    InputOutput::Variable* variable(const std::string& name, bool must_be_array,
                                    bool must_be_scalar = false) const;
    void check(int rc) const; // C-ABI status -> std::runtime_error
    CalcServer* _C;

  private:
    std::string _name;
    bool _once;
    Tool* _next = nullptr;
    int _id = -1;
    unsigned _n_iters = 0;
    double _elapsed_ms = 0.0;
};

/// type="kernel" (Kernel.cpp:301-352, 497-594): one registry kernel whose
/// arguments are bound to Variables by name
class Kernel : public Tool {
  public:
    Kernel(CalcServer* C, const std::string& name, const std::string& path,
           const std::string& entry, const std::string& n, bool once);
    void setup() override;
    const std::string& path() const { return _path; }
    const std::string& entry() const { return _entry; }
    const std::vector<InputOutput::Variable*>& arguments() const { return _vars; }
    int kernel_id() const { return _kid; }
    bool dependencies(std::vector<InputOutput::Variable*>& in,
                      std::vector<InputOutput::Variable*>& out) const override;
    /// Sweep fusion (aqc_fused_lookup): the leader launches the fused kernel for the
    /// whole group at its own position, the followers then have nothing left to do
    void fuse_lead(int fused_id, const std::vector<Kernel*>& group) { _fused_id = fused_id; _group = group; }
    void fuse_follow(Kernel* leader) { _leader = leader; }
    bool fused() const { return _fused_id >= 0 || _leader; }
    /// members of the fused group this kernel leads (empty: not a leader) / the leader it follows
    const std::vector<Kernel*>& group() const { return _group; }
    const Kernel* leader() const { return _leader; }
    int fusedId() const { return _fused_id; }
    /// true when the kernel walks neighbours (it has the link-list's head-of-cell argument)
    bool isSweep() const;
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
  private:
    size_t globalSize() const;
    int _fused_id = -1;
    std::vector<Kernel*> _group;
    Kernel* _leader = nullptr;
    std::string _path, _entry, _n;
    int _kid = -1;
    std::vector<InputOutput::Variable*> _vars;
    std::vector<int> _kinds;
    size_t _global = 0;
};

/// type="copy" (Copy.cpp:62-91)
class Copy : public Tool {
  public:
    Copy(CalcServer* C, const std::string& name, const std::string& in, const std::string& out, bool once)
      : Tool(C, name, once), _in_name(in), _out_name(out) {}
    void setup() override;
    bool dependencies(std::vector<InputOutput::Variable*>& in,
                      std::vector<InputOutput::Variable*>& out) const override
    {
        in.push_back(_in);
        out.push_back(_out);
        return true;
    }
    bool recordable(DeviceLoop&, std::string&) const override { return true; }
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
  private:
    std::string _in_name, _out_name;
    InputOutput::Variable *_in = nullptr, *_out = nullptr;
};

/// type="dummy": a named position in the pipeline (presets hang their tools before / after it)
class Dummy : public Tool {
  public:
    using Tool::Tool;
    bool recordable(DeviceLoop&, std::string&) const override { return true; }
};

/// type="set" (Set.cpp:197-226, Set.cl.in:32-47)
class Set : public Tool {
  public:
    Set(CalcServer* C, const std::string& name, const std::string& var, const std::string& value, bool once)
      : Tool(C, name, once), _var_name(var), _value(value) {}
    void setup() override;
    /// writes its array, reads scalars only (Set.cpp:60-75)
    bool dependencies(std::vector<InputOutput::Variable*>& in,
                      std::vector<InputOutput::Variable*>& out) const override
    {
        (void)in;
        out.push_back(_var);
        return true;
    }
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
  private:
    std::string _var_name, _value;
    InputOutput::Variable* _var = nullptr;
    bool _literal = false;        // OpenCL literal/macro: constant bytes in _data
    std::vector<char> _data;
};

/// Shared by set_scalar / if / while / assert (SetScalar.hpp ScalarExpression)
class ScalarExpression : public Tool {
  public:
    ScalarExpression(CalcServer* C, const std::string& name, const std::string& expr,
                     const std::string& type, bool once)
      : Tool(C, name, once), _expr(expr), _type(type) {}
    void setup() override;
  protected:
    void solve();                 // evaluates _expr into _value
    std::string _expr, _type;
    std::vector<char> _value;
};

/// type="set_scalar" (SetScalar.cpp:146-242)
class SetScalar : public ScalarExpression {
  public:
    SetScalar(CalcServer* C, const std::string& name, const std::string& var, const std::string& value, bool once)
      : ScalarExpression(C, name, value, "float", once), _var_name(var) {}
    void setup() override;
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void scalarOutputs(std::vector<InputOutput::Variable*>& out) const override { out.push_back(_var); }
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
  private:
    std::string _var_name;
    InputOutput::Variable* _var = nullptr;
};

/// type="assert" (Assert.cpp)
class Assert : public ScalarExpression {
  public:
    Assert(CalcServer* C, const std::string& name, const std::string& cond, bool once)
      : ScalarExpression(C, name, cond, "int", once) {}
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
};

/// type="if" / "while" (Conditional.cpp:40-144)
class Conditional : public ScalarExpression {
  public:
    Conditional(CalcServer* C, const std::string& name, const std::string& cond, bool once)
      : ScalarExpression(C, name, cond, "int", once) {}
    void setup() override;
    Tool* next_tool() override;
    int scope_modifier() const override { return 1; }
  protected:
    void _execute() override;
    bool _result = true;
    Tool* _ending_tool = nullptr;
};
/// `while`: the first pass runs tool by tool; when the body can be recorded (DeviceLoop::plan) the
/// rest of the loop runs as one CUDA graph from the first `end` on
class While : public Conditional {
  public:
    using Conditional::Conditional;
    ~While() override;
    /// called once every tool is set up and the sweeps are fused
    void planDeviceLoop();
    /// the matching `end` hands control back (End::next_tool)
    void reentry() { _reentry = true; }
    const DeviceLoop* deviceLoop() const { return _dev; }
    const std::string& hostReason() const { return _why; }
  protected:
    void _execute() override;
  private:
    DeviceLoop* _dev = nullptr;
    bool _reentry = false;
    std::string _why;
};
class If : public Conditional {
  public:
    using Conditional::Conditional;
    Tool* next_tool() override;
  protected:
    void _execute() override;
};
/// type="end" / "endif" (Conditional.cpp:155-181): jumps back to the opening tool
class End : public Tool {
  public:
    End(CalcServer* C, const std::string& name, bool once) : Tool(C, name, once) {}
    void setup() override;
    using Tool::next_tool;
    Tool* next_tool() override;
    int scope_modifier() const override { return -1; }
};

/// type="reduction" (Reduction.cpp:143-437)
class Reduction : public Tool {
  public:
    Reduction(CalcServer* C, const std::string& name, const std::string& in, const std::string& out,
              const std::string& operation, const std::string& null_val, bool once)
      : Tool(C, name, once), _in_name(in), _out_name(out), _operation(operation), _null(null_val) {}
    void setup() override;
    bool dependencies(std::vector<InputOutput::Variable*>& in,
                      std::vector<InputOutput::Variable*>& out) const override
    {
        in.push_back(_in);
        out.push_back(_out);
        return true;
    }
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void scalarOutputs(std::vector<InputOutput::Variable*>& out) const override { out.push_back(_out); }
    void record(DeviceLoop& L) override;
  protected:
    void _execute() override;
  private:
    std::string _in_name, _out_name, _operation, _null;
    InputOutput::Variable *_in = nullptr, *_out = nullptr;
    int _op = 0, _atype = 0;
    std::vector<char> _identity;
};

/// type="link-list" (LinkList.cpp:326-494).
/// `depends` (not a reference attribute; empty = the reference's behaviour) names the arrays whose
/// writes are the only way the input positions can change: while none of them is written through
/// the library a later execution finds its outputs (icell, ihoc, permutations) still valid and
/// does nothing -- the halo list of the slab pipelines inside the midpoint loop, whose input is
/// re-filled with the same positions in every sub-iteration.
class LinkList : public Tool {
  public:
    LinkList(CalcServer* C, const std::string& name, const InputOutput::ProblemSetup::Tool& t, bool once);
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _in_name, _min_name, _max_name, _ihoc_name, _icell_name, _ncells_name, _perm_name,
        _inv_name;
    bool _recompute;
    InputOutput::Variable *_in, *_min, *_max, *_ihoc, *_icell, *_ncells, *_perm, *_inv, *_N, *_support, *_h;
    std::string _depends_txt;
    std::vector<InputOutput::Variable*> _depends;
    int _watch = -1;
    const void* _built_in = nullptr;   // input pointer and length of the build the watch guards
    size_t _built_n = 0;
    uint64_t _skipped = 0;
  public:
    uint64_t skipped() const { return _skipped; }
};

/// type="radix-sort" / "sort" (RadixSort.cpp:129-303)
class RadixSort : public Tool {
  public:
    RadixSort(CalcServer* C, const std::string& name, const std::string& var, const std::string& perm,
              const std::string& inv, bool once)
      : Tool(C, name, once), _var_name(var), _perm_name(perm), _inv_name(inv) {}
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _var_name, _perm_name, _inv_name;
    InputOutput::Variable *_var = nullptr, *_perm = nullptr, *_inv = nullptr;
};

/// type="unsort" (UnSort.cl.in:30-42): out[perm[i]] = in[i]
class UnSort : public Tool {
  public:
    UnSort(CalcServer* C, const std::string& name, const std::string& in, const std::string& out,
           const std::string& perm, bool once)
      : Tool(C, name, once), _in_name(in), _out_name(out), _perm_name(perm) {}
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _in_name, _out_name, _perm_name;
    InputOutput::Variable *_in = nullptr, *_out = nullptr, *_perm = nullptr;
};

/// type="mpi-sync" (MPISync.cpp:183-232): particles whose `mask` names another
/// process travel there; what arrives is packed at the front of the same arrays.
/// The exchange runs over NCCL between device buffers (aqc_mpi_sync_ex).
/// `depends` (not a reference attribute; empty = the reference's behaviour) names the arrays the
/// mask content is a pure function of: while none of them is written, a later execution reuses
/// the sort permutation and the counts of the previous one and only gathers and exchanges --
/// the halo refresh inside the midpoint loop, where r is fixed.
class MPISync : public Tool {
  public:
    MPISync(CalcServer* C, const std::string& name, const std::string& mask,
            const std::string& fields, const std::string& procs, const std::string& depends, bool once)
      : Tool(C, name, once), _mask_name(mask), _fields_txt(fields), _procs_txt(procs),
        _depends_txt(depends) {}
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _mask_name, _fields_txt, _procs_txt, _depends_txt;
    InputOutput::Variable* _mask = nullptr;
    std::vector<InputOutput::Variable*> _fields, _depends;
    std::vector<unsigned> _procs;
    int _plan = -1;
};

/// type="mpi-allreduce" in="var" operation="min|max|sum" -- NOT a reference tool:
/// the reference has no collective, so dt / residuals are per process there
/// (SURVEY 5.8); multi-device cases of this repository add it after the
/// reductions whose result must be the same on every rank.
class MPIAllReduce : public Tool {
  public:
    MPIAllReduce(CalcServer* C, const std::string& name, const std::string& var,
                 const std::string& op, bool once)
      : Tool(C, name, once), _var_name(var), _op_txt(op) {}
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _var_name, _op_txt;
    InputOutput::Variable* _var = nullptr;
    int _op = 0, _type = 0;
    size_t _count = 1;
};

/// type="python" (Python.cpp:295-325).  The reference embeds CPython; this host hands the script to
/// the process that drives it through the runner registered with aqh_set_script_runner
/// (include/aquahost.h), which imports it and calls its main() with the `aquagpusph` module bound
/// to this simulation (aquagpusph_b200/pytool.py).  Without a runner the tool refuses to set up.
typedef int (*ScriptRunnerFn)(void* user, const char* tool_name, const char* script_path);
void setScriptRunner(ScriptRunnerFn fn, void* user);
class PythonTool : public Tool {
  public:
    PythonTool(CalcServer* C, const std::string& name, const std::string& path, bool once)
      : Tool(C, name, once), _path(path) {}
    void setup() override;
  protected:
    void _execute() override;
  private:
    std::string _path;
};

/// report_screen / report_file / report_dump / report_performance, and <Reports>
class Report : public Tool {
  public:
    Report(CalcServer* C, const std::string& name, const std::string& kind,
           const InputOutput::ProblemSetup::Tool& t, bool once);
    void setup() override;
    ~Report() override;
    bool recordable(DeviceLoop& L, std::string& why) const override;
    void record(DeviceLoop& L) override;
    const std::vector<InputOutput::Variable*>& fields() const { return _vars; }
  protected:
    void _execute() override;
  private:
    std::string _kind, _fields, _path;
    std::vector<InputOutput::Variable*> _vars;
    FILE* _f = nullptr;
};

/// End / print criteria (TimeManager.cpp:33-160)
class TimeManager {
  public:
    TimeManager(CalcServer* C, const InputOutput::ProblemSetup& sim_data);
    bool mustStop();
    bool mustPrintOutput();
    float time() const { return *_time; }
    float dt() const { return *_dt; }
    unsigned step() const { return *_step; }
    unsigned frame() const { return *_frame; }
  private:
    float *_time, *_dt, *_time_max;
    unsigned *_step, *_frame, *_steps_max, *_frames_max;
    float _output_time = 0.f, _output_fps = -1.f;
    unsigned _output_step = 0;
    int _output_ipf = -1;
};

/// The saver of one <Save> of a particles set (InputOutput/Particles.cpp:82-120, 243-323;
/// ASCII.cpp:240-333): device-side un-sort, download on the context's side stream into pinned
/// memory, file written by a thread that waits for the copy event (host/savers.cpp)
class ParticlesSaver {
  public:
    ParticlesSaver(CalcServer* C, size_t iset, size_t first, size_t n, const std::string& path,
                   const std::string& format, const std::string& fields);
    ~ParticlesSaver();
    /// Queue the download and start the writer; returns without waiting for either
    void save(float t);
    /// Particles::waitForSavers: block until the last file is complete (rethrows a writer error)
    void wait();
    size_t set() const { return _iset; }
    const std::string& file() const { return _file; }     ///< last file started
    const std::string& format() const { return _format; } ///< format actually written
    const std::string& fields() const { return _fields_txt; }
  private:
    struct Field;
    std::string nextFile();
    void print_file();
    CalcServer* _C;
    size_t _iset, _first, _n;
    std::string _path, _format, _ext, _fields_txt, _file, _error;
    char _sep = ',', _csep = ' '; // ASCII.hpp:83: print_file(',', ' ')
    unsigned _next_index = 0;
    float _time = 0.f;
    std::vector<std::unique_ptr<Field>> _fields;
    void* _event = nullptr;
    std::thread _writer;
};

/// The simulation: variables + tools on one device (CalcServer.cpp:139-621)
class CalcServer {
  public:
    CalcServer(InputOutput::ProblemSetup& sim_data, int device = -1, int mpi_rank = 0,
               int mpi_size = 1);
    ~CalcServer();
    /// Particle files -> device arrays (FileManager::load, Particles::loadDefault)
    void loadParticles();
    /// h check, per-set scalars, definitions, tool->setup() (CalcServer.cpp:1436-1534)
    void setup();
    /// Group neighbour sweeps that walk the same pairs (aqc_fused_lookup) when nothing
    /// between them in the pipeline reads their outputs or writes their inputs
    void planFusion();
    unsigned fused_groups() const { return _fused_groups; }
    /// `while` loops whose body runs as a CUDA graph (AQUA_DEVICE_LOOPS=0: none)
    void planDeviceLoops();
    unsigned device_loops() const { return _device_loops; }
    /// Run time steps until an output frame is due or the end criteria is met
    void update(TimeManager& t);
    /// Run exactly one pass over the pipeline (one time step)
    void step();

    InputOutput::Variables* variables() { return _vars.get(); }
    aqc_ctx* ctx() { return _ctx; }
    const std::vector<std::unique_ptr<Tool>>& tools() const { return _tools; }
    Tool* tool(size_t i) { return i < _tools.size() ? _tools[i].get() : nullptr; }
    InputOutput::ProblemSetup& sim_data() { return _sim_data; }
    int dims() const { return _sim_data.dims; }
    int mpi_rank() const { return _mpi_rank; }
    int mpi_size() const { return _mpi_size; }
    const std::vector<std::pair<std::string, std::string>>& definitions() const { return _defs; }
    bool lapMorris() const { return _lap_morris; }
    /// Download an array in the ORIGINAL particle order (CalcServer::getUnsortedMem,
    /// CalcServer.cpp:772-830); `out` must hold length*typesize bytes
    void getUnsortedMem(const std::string& var, void* out);
    void download(const std::string& var, void* out);
    void upload(const std::string& var, const void* in);
    /// FileManager::save (FileManager.cpp:146-155): every set's <Save> file (asynchronously, see
    /// ParticlesSaver) and the AQUAgpusph.save.N.xml state file the run can be resumed from
    void save(float t);
    /// FileManager::waitForSavers (FileManager.cpp:157-164)
    void waitForSavers();
    const std::string& checkpoint_file() const { return _checkpoint_file; }
    const std::vector<std::unique_ptr<ParticlesSaver>>& savers() const { return _savers; }
    uint64_t steps_done() const { return _steps; }
    /// Join the NCCL communicator of the run (id = 128 bytes made by rank 0 with
    /// aqc_comm_unique_id and distributed by the launcher); replaces MPI_Init
    void commInit(const void* unique_id);

  private:
    void buildDefinitions();
    Tool* makeTool(const InputOutput::ProblemSetup::Tool& t);
    InputOutput::ProblemSetup& _sim_data;
    aqc_ctx* _ctx = nullptr;
    std::unique_ptr<InputOutput::Variables> _vars;
    std::vector<std::unique_ptr<Tool>> _tools;
    std::vector<std::pair<std::string, std::string>> _defs; // name -> value as "-D" text
    bool _lap_morris = false; // __LAP_FORMULATION__ = __LAP_MORRIS__: cfd/Interactions.cl is never fused
    int _mpi_rank, _mpi_size;
    uint64_t _steps = 0;
    unsigned _fused_groups = 0;
    unsigned _device_loops = 0;
    void* _unsort_scratch = nullptr;
    size_t _unsort_cap = 0;
    void writeCheckpoint();
    std::vector<std::unique_ptr<ParticlesSaver>> _savers;
    std::string _checkpoint_file;
};

} // namespace CalcServer
} // namespace Aqua
