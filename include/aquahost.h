/* aquahost.h -- C entry points of libaquahost.so, the C++ host that keeps the
 * reference's XML problem/tool API (aquagpusph/main.cpp:108-200,
 * FileManager.cpp:60-143, CalcServer.cpp:592-621) on top of libaquacuda.so.
 *
 * It is what `AQUAgpusph -i Main.xml -d 3` does, split so that another process
 * (bench.py, the tests) can drive it: load a case, step it, move particle
 * arrays in and out with HOST buffers.  Every function returns 0 on success and
 * <0 on error (message in aqh_last_error()); no C++ exception crosses it. */
#ifndef AQUAHOST_H
#define AQUAHOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aqh_sim aqh_sim;

const char* aqh_last_error(void);
void aqh_set_log_level(int level); /* 0 debug .. 3 error (ArgumentsManager.cpp:98-110) */

/* FileManager::load + new CalcServer + setup: parse xml_path (dims = 2|3, the -d
 * flag), create the device context (device < 0: take it from <Device>), register
 * variables/tools, load the particle files and set every tool up.  root_path is the
 * folder holding "resources/" (may be NULL: AQUAGPUSPH_ROOT, then <RootPath>). */
int aqh_load(const char* xml_path, int dims, int device, const char* root_path, int mpi_rank,
             int mpi_size, aqh_sim** out);
/* XML front-end only (no device needed): parse and resolve the includes, presets
 * and tool placement; the result serves aqh_write_resolved / aqh_n_tools /
 * aqh_tool_name / aqh_tool_type. */
int aqh_parse(const char* xml_path, int dims, const char* root_path, aqh_sim** out);
/* Multi-device runs (one process per GPU; replaces Aqua::MPI::init, main.cpp:112):
 * rank 0 makes the 128-byte id, the launcher distributes it (torch.distributed
 * broadcast, a file ...), every rank loaded with mpi_size > 1 joins before stepping. */
int aqh_comm_unique_id(void* id_out_128_bytes);
int aqh_comm_init(aqh_sim* sim, const void* unique_id);
void aqh_destroy(aqh_sim* sim);
/* State::write: the resolved problem as one flat XML (checkpoint format) */
int aqh_write_resolved(aqh_sim* sim, const char* path);

/* pipeline introspection */
int aqh_n_tools(aqh_sim* sim);
const char* aqh_tool_name(aqh_sim* sim, int i);
const char* aqh_tool_type(aqh_sim* sim, int i); /* XML type attribute */
double aqh_tool_elapsed_ms(aqh_sim* sim, int i); /* accumulated host time */
unsigned aqh_tool_used_times(aqh_sim* sim, int i);

/* CalcServer::update split in steps: run n passes over the pipeline */
int aqh_step(aqh_sim* sim, int n);
/* main.cpp:162-181: run until the end criteria -- update until an output frame is due
 * (TimeManager::mustPrintOutput), save it (aqh_save), ... -- then wait for the writers */
int aqh_run(aqh_sim* sim);
/* FileManager::save (FileManager.cpp:146-155, Particles.cpp:82-120, ASCII.cpp:240-333,
 * State.cpp:195-226 + 1517-1908): every <Save> of every particles set goes to its next numbered file
 * (un-sorted on the device, downloaded on a side stream, written by a thread: the call returns at
 * once) and the state file AQUAgpusph.save.N.xml is rewritten: all variables with their current
 * values, the tools, the timing options and a <Load> of the files just started -- loading it with
 * aqh_load resumes the run.  aqh_wait_savers = FileManager::waitForSavers (blocks until the files
 * are complete; a writer's error surfaces here). */
int aqh_save(aqh_sim* sim);
int aqh_wait_savers(aqh_sim* sim);
const char* aqh_checkpoint_file(aqh_sim* sim); /* "" before the first aqh_save */
int aqh_n_savers(aqh_sim* sim);
const char* aqh_saver_file(aqh_sim* sim, int i); /* last file of saver i */
int aqh_sync(aqh_sim* sim);
uint64_t aqh_launch_count(aqh_sim* sim); /* CUDA kernels launched so far */
unsigned aqh_fused_groups(aqh_sim* sim); /* sweep groups the planner fused (0 with AQUA_NO_FUSION) */
void* aqh_cuda_ctx(aqh_sim* sim);         /* the aqc_ctx* underneath */

/* Variables + Tokenizer without a device (Variable.cpp:1321-1435): register the
 * scalar variables listed in `decls` ("type name=value;type name=value;...", in
 * order, values are expressions that may use the earlier names), then evaluate
 * `expr` as a value of `type` into out.  What `set_scalar` does on the host. */
int aqh_eval(int dims, const char* decls, const char* type, const char* expr, void* out,
             size_t bytes);

/* The same value computed the way a recorded loop computes it: `expr` is compiled
 * into a stack program of include/aquasvm.h over a table holding every declared 32-bit scalar and
 * run by the interpreter the device kernel is built from -- on the host.  The CPU tests pin it to
 * aqh_eval bit for bit. */
int aqh_eval_svm(int dims, const char* decls, const char* type, const char* expr, void* out,
                 size_t bytes);
/* `while` loops of the pipeline that run as a CUDA graph from their second pass on
 * (host/devloop.hpp; AQUA_DEVICE_LOOPS=0: none), how many times they ran on the device and the
 * passes they made there; why the `while` at tool index i stays on the host ("" when it does not,
 * NULL when tool i is not a `while`) */
unsigned aqh_device_loops(aqh_sim* sim);
int aqh_device_loop_stats(aqh_sim* sim, uint64_t* runs, uint64_t* iterations);
const char* aqh_loop_host_reason(aqh_sim* sim, int i);
/* graph nodes of the recorded bodies, host milliseconds the last recording and the last
 * cudaGraphInstantiate took (summed over the loops): the per-step host cost of the device loops */
/* tools of the device loops' bodies that run on the second stream (aqc_lane_*: their arrays do not
 * meet those of the tools running meanwhile on the first; AQUA_DEVICE_LANES=0: none) */
unsigned aqh_device_loop_branch_tools(aqh_sim* sim);
/* The schedule behind it as a pure function (no device): n tools in pipeline order; tool k reads the arrays
 * r_var[r_off[k] .. r_off[k+1]) (ids) in the rows r_rows[..] (AQC_ROWS_* masks) and writes w_var / w_rows
 * likewise; cost[k] in element-wise-kernel units; flags[k]: 1 = must stay on lane 0, 2 = unknown dependencies
 * (conflicts with everything), 4 = launches nothing.  Out: lane_out[k] in {0, 1}; wait_out[k] = the tool of
 * the other lane whose event tool k waits for (-1: none); marked_out[k] = an event is recorded behind tool k.
 * *last_lane1_out = the last tool on lane 1, joined at the end of the pass (-1: a single lane).
 * tests/test_host_cpu.py checks on random instances that every conflicting pair is ordered by lane order
 * and event waits. */
int aqh_lane_schedule(int n, const int* r_off, const int* r_var, const unsigned* r_rows, const int* w_off,
                      const int* w_var, const unsigned* w_rows, const double* cost, const unsigned char* flags,
                      double gain, int* lane_out, int* wait_out, unsigned char* marked_out, int* last_lane1_out);
int aqh_device_loop_timing(aqh_sim* sim, int* body_nodes, double* record_ms, double* instantiate_ms);

/* type="python" tools (aquagpusph/CalcServer/Python.cpp:295-325).  The reference embeds CPython in
 * its host; this host calls `fn(user, tool name, script path)` instead, once per execution of the
 * tool and after a device sync, and the driving process runs the script's main() with get / set
 * bound to aqh_scalar_get / aqh_scalar_set / aqh_array_download / aqh_array_upload
 * (aquagpusph_b200/pytool.py).  Non-zero = "Python execution error" (or main() returned False):
 * the step fails like Python.cpp:304-318.  Process-wide; register before aqh_load -- a problem
 * with a python tool and no runner fails at set-up. */
typedef int (*aqh_script_fn)(void* user, const char* tool_name, const char* script_path);
void aqh_set_script_runner(aqh_script_fn fn, void* user);

/* variables */
/* the reference type string of a variable ("float", "vec", "unsigned int*", ...), NULL when it is
 * not declared; valid until the next call on this thread */
const char* aqh_variable_type(aqh_sim* sim, const char* name);
int aqh_scalar_get(aqh_sim* sim, const char* name, void* out, size_t bytes);
int aqh_scalar_set(aqh_sim* sim, const char* name, const char* expression);
int aqh_array_info(aqh_sim* sim, const char* name, size_t* length, size_t* elem_bytes);
/* unsorted != 0: original particle order (CalcServer::getUnsortedMem) */
int aqh_array_download(aqh_sim* sim, const char* name, void* host_out, int unsorted);
int aqh_array_upload(aqh_sim* sim, const char* name, const void* host_in);
void* aqh_array_devptr(aqh_sim* sim, const char* name);

#ifdef __cplusplus
}
#endif
#endif
