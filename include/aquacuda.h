/* aquacuda.h -- C-ABI of libaquacuda.so: the B200 (sm_100a) device layer that
 * replaces the OpenCL back-end of AQUAgpusph's CalcServer tools.
 *
 * The reference has no FFI; its "device boundary" is the OpenCL host API called
 * from the tool classes (aquagpusph/CalcServer/{Kernel,LinkList,RadixSort,
 * Reduction,Set,Copy,UnSort,MPISync}.cpp).  Every entry point below names the
 * reference interface it stands in for.  A maintainer of the reference would
 * call these from Tool::_execute() overrides (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *  - plain C: opaque context, raw device pointers (void*), sizes; no C++/torch
 *    types.  Every function returns 0 on success, <0 on error and stores a
 *    message retrievable with aqc_last_error() (the reference's equivalent is
 *    CHECK_OCL_OR_THROW, CalcServer.hpp:44-49: the C++ side turns non-zero into
 *    std::runtime_error at the same places).
 *  - all work is enqueued on the context's CUDA stream (in order); nothing
 *    blocks the host unless stated ("syncs").
 *  - "usize" is 32 bit (the reference default when <Device> has no addr_bits,
 *    State.cpp:499-502); vec = float4 in 3-D, float2 in 2-D
 *    (resources/Scripts/types/3D.h:23-31, 2D.h:23-31); matrix = float16/float4.
 *  - there is NO CPU fallback: without a CUDA device aqc_ctx_create fails.
 */
#ifndef AQUACUDA_H
#define AQUACUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aqc_ctx aqc_ctx;
typedef uint32_t aqc_usize;

#define AQC_OK 0
#define AQC_ERR_CUDA (-1)     /* a CUDA runtime call failed */
#define AQC_ERR_ARG (-2)      /* invalid argument */
#define AQC_ERR_NOKERNEL (-3) /* (script, entry) not in the registry */
#define AQC_ERR_NCCL (-4)
#define AQC_ERR_STATE (-5)

/* ---- context: replaces CalcServer::setupOpenCL (CalcServer.cpp:858-886) and
 * the command-queue pool (CalcServer.cpp:674-705) ------------------------- */
int aqc_ctx_create(int device, aqc_ctx** out);
void aqc_ctx_destroy(aqc_ctx* ctx);
const char* aqc_last_error(const aqc_ctx* ctx);
/* Run on an externally owned cudaStream_t (e.g. the caller's). NULL = own. */
int aqc_set_stream(aqc_ctx* ctx, void* cuda_stream);
void* aqc_get_stream(aqc_ctx* ctx);
int aqc_sync(aqc_ctx* ctx); /* clFinish */
/* number of kernels this library launched on ctx since creation */
uint64_t aqc_launch_count(const aqc_ctx* ctx);
int aqc_device_sm_count(const aqc_ctx* ctx);

/* ---- "-D" definitions baked into every OpenCL kernel by the reference
 * (CalcServer.cpp:240-265, basic.xml:119-123, cfd.xml:52-54).  Here they are
 * constant-bank parameters of the context; H/CONW/CONF must already be the
 * 6-significant-digit values the reference would print (see
 * aqc_define_round6). ------------------------------------------------------ */
typedef struct {
    int dims;      /* 2 | 3  (-DHAVE_2D / -DHAVE_3D, Tool.cpp:330-333) */
    float H;       /* -DH */
    float CONW;    /* -DCONW */
    float CONF;    /* -DCONF */
    float SUPPORT; /* -DSUPPORT (2.f) */
    float DIMS;    /* -DDIMS (evaluated => float literal) */
} aqc_defs;
float aqc_define_round6(float value); /* CalcServer.cpp:245-257 "%#G" + "f" */
int aqc_set_defs(aqc_ctx* ctx, const aqc_defs* defs);
/* Any other "-Dname=value" of the problem (CalcServer.cpp:240-265), as text.  The
 * ones the CUDA kernels honour: __DR_FACTOR__, __MIN_BOUND_DIST__ (BIe/ElasticBounce.cl:
 * 31-36, PST.cl:31-33), TSCHEME_ADAMS_BASHFORTH_STEPS (basic/time_scheme/adam_bashforth.cl:66-68),
 * plus DIMS/H/CONW/CONF/SUPPORT (same as aqc_set_defs).  Returns 0
 * when the definition was consumed, 1 when it is not used by any CUDA kernel, and <0
 * when it selects something this build does not provide (KERNEL_NAME other than
 * Wendland, __LAP_FORMULATION__ other than __LAP_MONAGHAN__ = 1). */
int aqc_set_define(aqc_ctx* ctx, const char* name, const char* value);

/* ---- device memory: ArrayVariable storage (Variable.cpp:1739-1822) ------- */
int aqc_alloc(aqc_ctx* ctx, size_t bytes, void** dptr);
int aqc_free(aqc_ctx* ctx, void* dptr);
int aqc_host_alloc(aqc_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int aqc_host_free(aqc_ctx* ctx, void* hptr);
/* clEnqueueWriteBuffer / ReadBuffer / CopyBuffer (Copy.cpp:62-91); blocking!=0 syncs */
int aqc_memcpy_h2d(aqc_ctx* ctx, void* dst, const void* src, size_t bytes, int blocking);
int aqc_memcpy_d2h(aqc_ctx* ctx, void* dst, const void* src, size_t bytes, int blocking);
int aqc_memcpy_d2d(aqc_ctx* ctx, void* dst, const void* src, size_t bytes);

/* Asynchronous device->host copy on the context's SIDE stream: the reference's savers download on
 * a command queue of their own (Particles.cpp:243-323: C->command_queue(cmd_queue_new), events joined
 * by a marker) so that writing files overlaps the next time steps.  aqc_side_fork makes the side
 * stream wait for everything queued so far on the main stream; aqc_memcpy_d2h_side queues one copy
 * there (dst: pinned memory of aqc_host_alloc); aqc_side_record records `ev` (aqc_event_create)
 * behind the copies: aqc_side_wait(ev) from any thread then waits for the data only.  The main
 * stream never waits for the side stream: a source buffer must not be rewritten before `ev`. */
int aqc_side_fork(aqc_ctx* ctx);
int aqc_memcpy_d2h_side(aqc_ctx* ctx, void* dst, const void* src, size_t bytes);
int aqc_side_record(aqc_ctx* ctx, void* ev);
/* Wait for `ev` from ANY host thread (the savers' writer threads): selects the context's device for
 * the calling thread and blocks on the event alone -- no communicator polling, no context state. */
int aqc_side_wait(aqc_ctx* ctx, void* ev);

/* ---- Set tool (Set.cpp:197-226, Set.cl.in:32-47): fill n elements of
 * elem_bytes (4, 8, 16 or 64) with the value at *value ---------------------- */
int aqc_fill(aqc_ctx* ctx, void* dptr, size_t n, size_t elem_bytes, const void* value);

/* ---- LinkList tool (LinkList.cpp:326-494; kernels LinkList.cl.in:32-113).
 * r: N vec; icell, perm (= id_unsorted), inv_perm (= id_sorted): N usize.
 * recompute_grid != 0: r_min/r_max are reduced from r (LinkList.cpp:338-356),
 * else the host values passed in are used.  rmin/rmax/ncells are HOST arrays of
 * 4 entries (in/out).  *ihoc / *ihoc_capacity (elements) are in/out: when
 * n_cells.w exceeds the capacity the library frees *ihoc, allocates a larger
 * buffer and returns it (LinkList::allocate, LinkList.cpp:234-271).
 * Syncs once (the reference blocks at the same place, LinkList.cpp:350-356). */
int aqc_linklist_build(aqc_ctx* ctx, const void* r, aqc_usize N, int dims,
                       float support, float h, int recompute_grid,
                       float rmin[4], float rmax[4], aqc_usize ncells[4],
                       aqc_usize* icell, aqc_usize** ihoc, size_t* ihoc_capacity,
                       aqc_usize* perm, aqc_usize* inv_perm);

/* ---- RadixSort tool (RadixSort.cpp:129-303): stable ascending sort of n usize
 * keys in place; perm[k] = input index of the k-th output
 * (RadixSort.cl.in:295-296), inv_perm[perm[k]] = k (:313-323).  key_max is an
 * upper bound of the key values (n_cells.w for "icell", else 0 => full 32 bit;
 * RadixSort.cpp:152-176).  perm / inv_perm may be NULL. */
int aqc_radix_sort(aqc_ctx* ctx, aqc_usize* keys, aqc_usize n, aqc_usize key_max,
                   aqc_usize* perm, aqc_usize* inv_perm);

/* ---- UnSort tool (UnSort.cl.in:30-42) and the particle permutation of
 * basic/Sort.cl:57-124: for f < nfields, dst[f][idx[i]] = src[f][i] --------- */
int aqc_scatter_fields(aqc_ctx* ctx, const aqc_usize* idx, aqc_usize N, int nfields,
                       const void* const* src, void* const* dst,
                       const size_t* elem_bytes);

/* ---- Reduction tool (Reduction.cpp:143-258; Reduction.cl.in:35-66).
 * op: see enum; type: see enum.  The result is written to out_dev (device, may
 * be NULL) and, when out_host != NULL, copied there (that variant syncs).
 * Sums are evaluated in a fixed order (run-to-run deterministic). */
enum { AQC_OP_SUM = 0, AQC_OP_MIN = 1, AQC_OP_MAX = 2 };
enum { AQC_T_F32 = 0, AQC_T_U32 = 1, AQC_T_I32 = 2, AQC_T_VEC2 = 3, AQC_T_VEC4 = 4 };
int aqc_reduce(aqc_ctx* ctx, int op, int type, const void* in, size_t n,
               void* out_dev, void* out_host);

/* ---- Kernel tool (Kernel.cpp:301-352, 497-556).  The reference compiles an
 * OpenCL script and reflects the argument NAMES with clGetKernelArgInfo; here
 * the hot-path scripts are pre-built CUDA kernels kept in a registry keyed by
 * the script path (as written in the presets, e.g. "cfd/Interactions.cl" --
 * any leading directories up to "Scripts/" are ignored) and entry point. ----- */
/* ARRAY_RO: declared without const in the reference's script but never written by the kernel
 * (e.g. iset, imove, rho of basic/EOS.cl:57-73): a read for every dependency purpose */
enum { AQC_ARG_ARRAY_IN = 0, AQC_ARG_ARRAY_OUT = 1, AQC_ARG_SCALAR = 2, AQC_ARG_ARRAY_RO = 3 };
typedef struct {
    const char* name; /* Variable name to bind (Kernel.cpp:497-556) */
    const char* type; /* reference type string: "vec*", "float", "usize", "svec4", ... */
    int kind;         /* AQC_ARG_* ; ARRAY_OUT = non-const __global pointer */
} aqc_arg_info;
/* returns kernel id >= 0, or AQC_ERR_NOKERNEL */
int aqc_kernel_lookup(const char* script_path, const char* entry, int dims);
int aqc_kernel_count(void);
const char* aqc_kernel_name(int kernel_id); /* "cfd/Interactions.cl::entry" */
int aqc_kernel_nargs(int kernel_id);
const aqc_arg_info* aqc_kernel_args(int kernel_id);
/* clSetKernelArg + clEnqueueNDRangeKernel (Kernel.cpp:324-352): args[k] is the
 * device pointer for array arguments, or a HOST pointer to the scalar value
 * (float, uint, vec = 2/4 floats, svec4 = 4 uints) for scalar arguments, in
 * registry order.  n = global work size (Kernel.cpp:558-594). */
int aqc_launch(aqc_ctx* ctx, int kernel_id, size_t n, void* const* args, int nargs);
/* A script that is NOT in the registry (case-local *.cl, families without hand-written kernels): what
 * Kernel::make does with every script in the reference (Kernel.cpp:354-420: read the source, "-I<script
 * folder> -I<base path>", the problem's -D definitions, clBuildProgram; :497-556 argument names and
 * qualifiers from clGetKernelArgInfo).  The script and the headers it includes are read where they lie,
 * compiled for sm_100a by NVRTC behind a dialect header, and entered in the registry: the returned id works
 * with aqc_kernel_nargs / aqc_kernel_args / aqc_launch.  defines: "-DNAME=VALUE" strings of
 * CalcServer.cpp:240-265.  Neighbour loops (BEGIN_NEIGHS) run as the script wrote them, one particle
 * per thread; registry kernels are never replaced by this path.  < 0: error (libnvrtc missing, compile error:
 * aqc_last_error holds the log). */
int aqc_script_compile(aqc_ctx* ctx, const char* path, const char* entry, int dims, const char* base_path,
                       const char* const* defines, int ndefines);
/* the same up to the cubin, without a device and without registering anything: returns the number of
 * arguments (their "type name;" list goes to `log`), or < 0 with the compiler's message in `log` */
int aqc_script_check(const char* path, const char* entry, int dims, const char* base_path,
                     const char* const* defines, int ndefines, char* log, size_t log_bytes);

/* ---- fusion of neighbour sweeps.  The reference launches one sweep per script
 * kernel (9 per midpoint sub-iteration in the 3-D dam break); sweeps
 * that walk the same (i, j) pairs can share the candidate filter and the pair
 * geometry.  kernel_ids: n kernel ids in pipeline order.  Returns a fused id, or
 * AQC_ERR_NOKERNEL when this set is not fused (the caller then launches the
 * members one by one).  The arguments of aqc_launch_fused are the members'
 * argument lists concatenated in the same order; the results equal those of launching
 * the members one after the other up to FMA contraction (a few ulp).  It is the CALLER's business to
 * check that nothing between the members' positions in its pipeline reads the
 * outputs or writes the inputs (the C++ host does, calcserver.cpp). */
int aqc_fused_lookup(const int* kernel_ids, int n, int dims);
/* 1 when the list is the beginning of (or all of) some fused set, else 0 */
int aqc_fused_prefix(const int* kernel_ids, int n, int dims);
/* particle classes (imove) whose ROWS a kernel writes / a fused sweep reads besides the
 * positions; lets a caller see that cfd/Sensors.cl (sensor rows of u, rho, p) does not
 * disturb a sweep over fluid pairs */
enum { AQC_ROWS_FLUID = 1, AQC_ROWS_SENSOR = 2, AQC_ROWS_BOUNDARY = 4, AQC_ROWS_ANY = 7 };
int aqc_kernel_write_rows(int kernel_id);
/* ... and the rows whose VALUES a stand-alone kernel uses in the arrays other than the positions
 * (a kernel may still load rows it then discards by their class: a concurrent writer of those rows
 * changes nothing it computes).  AQC_ROWS_ANY unless the script says otherwise. */
int aqc_kernel_read_rows(int kernel_id);
int aqc_fused_read_rows(int fused_id);
int aqc_launch_fused(aqc_ctx* ctx, int fused_id, void* const* args, int nargs);
/* Neighbour-sweep engine of the kernel-support sweeps (both are CUDA):
 * 3 = CTA-shared tiles with deferred pair bodies (default), 2 = per-warp tiles, bodies in the
 * reference's visiting order.  Initial value from AQC_SWEEP_ENGINE.  Returns the engine in
 * use after the call; engine = 0 only queries, engine < 0 returns to the automatic choice (the
 * environment, else v3 except for 2-D problems below ~1 M particles, where the per-warp engine
 * is faster).  For A/B measurements and order-sensitive tests. */
int aqc_sweep_engine_select(int engine);

/* ---- pair-mask cache of the neighbour sweeps.  Every BEGIN_NEIGHS/END_NEIGHS kernel of the
 * reference re-walks the 27 (9) cells and re-tests every candidate (types/3D.h:197-219), although
 * between two link-list builds most sweeps see the same positions: the midpoint scheme keeps r
 * fixed over its sub-iterations (basic/time_scheme/midpoint.cl:93-111), so MLS and the three
 * fluid / lapp_corr passes of a step find the same neighbours.  With the cache enabled the
 * first such sweep after a change builds the hit masks of the candidate filter (device memory:
 * ~1 KB per particle in 3-D) and the following ones read them instead of filtering; every pair
 * is still re-tested exactly, so the results are bit-identical with and without the cache.
 * The cache is dropped whenever r, imove, icell or ihoc is written THROUGH THIS LIBRARY (copies,
 * fills, kernel outputs, link-list, sort, scatter, mpi-sync); a caller that writes them by other
 * means (its own kernels) must call aqc_pairs_cache_invalidate.  Off by default. ------------- */
int aqc_pairs_cache_enable(aqc_ctx* ctx, int on);
int aqc_pairs_cache_invalidate(aqc_ctx* ctx);
/* builds / sweeps served so far, bytes of device memory held (any pointer may be NULL) */
int aqc_pairs_cache_stats(const aqc_ctx* ctx, uint64_t* builds, uint64_t* hits, uint64_t* bytes);
/* ... and of the second cache, the one of the REMOTE (halo) sweeps of cfd/MPI.cl / aqua/MPIdeltaSPH.cl: their
 * neighbour lists are built once per halo link-list and read by every remote sweep until the local
 * geometry or the halo list changes (AQC_REMOTE_LISTS=0: the remote sweeps filter every time) */
int aqc_pairs_cache_stats_remote(const aqc_ctx* ctx, uint64_t* builds, uint64_t* hits, uint64_t* bytes);

/* ---- write watches: a set of device ranges that turns dirty as soon as one of them is written
 * through this library (the mechanism behind the pair cache and the mpi-sync plans, for callers).
 * The host's link-list tool uses one to skip the re-build of a list whose positions cannot have
 * changed (`depends`).  A new watch is dirty; aqc_watch_reset arms it. ------ */
int aqc_watch_create(aqc_ctx* ctx); /* returns a watch id >= 0 */
int aqc_watch_dirty(const aqc_ctx* ctx, int watch);
int aqc_watch_reset(aqc_ctx* ctx, int watch, int n, const void* const* ptrs, const size_t* bytes);

/* ---- multi-device: one process per GPU, NCCL over NVLink.  Replaces the MPI
 * rank/size queries and wrappers (AuxiliarMethods.cpp:388-516) and the MPISync
 * tool (MPISync.cpp:183-232; kernels MPISync.cl.in:31-80), whose host-staged
 * clEnqueueReadBuffer -> MPI_Isend / MPI_Recv -> clEnqueueWriteBuffer per field
 * (MPISync.cpp:564-638, 932-1052) becomes grouped ncclSend/ncclRecv between
 * device buffers.  NCCL is dlopen'ed on first use. ---------------------------- */
#define AQC_UNIQUE_ID_BYTES 128
/* rank 0 creates the id (ncclGetUniqueId); the caller distributes the 128 bytes
 * (torch.distributed broadcast, a file, MPI ...) and every rank calls init */
int aqc_comm_unique_id(void* id_out);
int aqc_comm_init(aqc_ctx* ctx, int rank, int size, const void* unique_id);
int aqc_comm_destroy(aqc_ctx* ctx);
int aqc_comm_rank(const aqc_ctx* ctx);
int aqc_comm_size(const aqc_ctx* ctx);
/* MPISync::_execute.  mask: n usize, the destination process of every element
 * (elements with mask == own rank stay).  The mask is sorted (stable), the fields
 * are gathered in that order, the block bound to each process of `procs`
 * (NULL/0: every other rank) is sent, and what the peers send is packed at the
 * FRONT of the same field arrays in process order; on return mask[k] = sender
 * rank over the received blocks and = own rank elsewhere.  *n_received (host,
 * may be NULL) = number of elements received.  Syncs once (counts).  With one
 * rank nothing happens, like MPISync.cpp:186-187. */
int aqc_mpi_sync(aqc_ctx* ctx, aqc_usize* mask, aqc_usize n, int nfields, void* const* fields,
                 const size_t* elem_bytes, int nprocs, const unsigned* procs,
                 aqc_usize* n_received);
/* The same with a PLAN (one slot per mpi-sync tool, aqc_mpi_sync_plan): the caller names the device
 * ranges the mask content is a pure function of (`dep_ptrs`/`dep_bytes`: e.g. r and imove for the
 * plane masks of cfd/MPI/planes.cl).  A later call with the same plan, mask, fields and processes
 * finds the sort permutation, the counts and the receive layout of the previous one still valid
 * unless one of those ranges was written through this library in between, and then only gathers
 * and exchanges: no sort, no count all-gather, no host synchronisation.  The reference sorts and
 * exchanges counts on every call (MPISync.cpp:183-232); inside the midpoint loop r is fixed
 * (basic/time_scheme/midpoint.cl:93-111), so the halo of every sub-iteration but the first
 * travels on the first one's plan.  ndeps = 0 or plan < 0: every call is a full one.
 * AQC_MPI_VERIFY=1 checks every reuse against a copy of the mask (tests).
 * Failure behaviour of every collective entry point: while a communicator is live, waits for the
 * device are bounded (AQC_COMM_TIMEOUT_S, default 60 s); a local device fault, an NCCL error or an
 * expired wait aborts the communicator (ncclCommAbort) and the call fails, after which every
 * collective call of this context fails at once -- a dead rank ends the job instead of leaving
 * its peers inside ncclRecv. */
int aqc_mpi_sync_plan(aqc_ctx* ctx); /* returns a new plan id >= 0 */
int aqc_mpi_sync_ex(aqc_ctx* ctx, int plan, aqc_usize* mask, aqc_usize n, int nfields,
                    void* const* fields, const size_t* elem_bytes, int nprocs, const unsigned* procs,
                    aqc_usize* n_received, int ndeps, const void* const* dep_ptrs,
                    const size_t* dep_bytes);
/* full (sorted + counted) and reused executions of a plan so far */
int aqc_mpi_sync_stats(const aqc_ctx* ctx, int plan, uint64_t* full, uint64_t* reused);
/* ADDITIONS to the reference, which has no collective: element-wise
 * all-reduce (AQC_OP_*, AQC_T_*) of a device array in place / of a small host
 * value (<= 64 bytes; syncs).  aqc_linklist_build all-reduces r_min / r_max by
 * itself when a communicator exists, so every rank hashes on one global grid. */
int aqc_allreduce(aqc_ctx* ctx, int op, int type, void* dev_inout, size_t count);
int aqc_allreduce_host(aqc_ctx* ctx, int op, int type, void* host_inout, size_t count);

/* ---- lanes.  The reference runs tools whose dependencies allow it on different OpenCL command
 * queues (CalcServer.cpp:674-705: the queue pool; Tool.cpp:405-444: the events of the variables a tool
 * reads and writes are what it waits for).  A context here is ONE in-order stream, plus a second one
 * on request: aqc_lane_select(ctx, 1) sends every following call of the context to the BRANCH lane
 * (own stream, lower priority, own sweep scratch) until aqc_lane_select(ctx, 0).  Ordering between
 * the lanes is the caller's business: aqc_lane_event records an event (created when *ev is NULL) behind
 * what was queued on the lane in use, aqc_lane_wait makes the lane in use wait for one.  Both work
 * while a loop body records (the branch lane joins the capture through its first wait and must be
 * waited for by lane 0 before aqc_loop_end).  Reductions, scalar programs, link-list, sort, mpi-sync
 * and every call that synchronises belong on lane 0. ------------------------------------------- */
int aqc_lane_select(aqc_ctx* ctx, int lane);
int aqc_lane_event(aqc_ctx* ctx, void** ev);
int aqc_lane_wait(aqc_ctx* ctx, void* ev);

/* ---- device-side loops.  The reference evaluates `while` conditions and
 * set_scalar expressions on the host behind the events of the variables they read
 * (Conditional.cpp:85-96, SetScalar.cpp:146-195) and reads every reduction back before the host
 * can decide what to enqueue next (Reduction.cpp:205-258).  A loop object records ONE pass over a
 * loop body -- every call made on the context between aqc_loop_begin and aqc_loop_end is
 * captured instead of executed -- as the body of a CUDA graph WHILE node; the scalar tools of the
 * body are small stack programs (include/aquasvm.h) over a table of typed scalars in device
 * memory, run by a one-thread kernel, and the program holding AQS_SETCOND sets the node's
 * condition.  aqc_loop_run uploads the table, launches the graph (entry program -> while (cond)
 * { body }), downloads header, table and report history, and syncs ONCE, however many
 * iterations ran.  Only calls that neither synchronise nor read back may be made while a loop
 * records (kernel launches, fills, device copies, aqc_reduce with out_host == NULL): anything
 * else fails the recording -- aqc_loop_end then returns an error, nothing has been executed, the
 * pair caches and watches are invalidated, and the caller runs the loop its old way. ---------- */
typedef struct aqc_loop aqc_loop;
struct aqs_op;     /* aquasvm.h */
struct aqs_header; /* aquasvm.h */
/* table_bytes: size of the scalar table (multiple of 16); hist_rows: report snapshots kept per run;
 * max_ops: capacity of the program arena */
int aqc_loop_create(aqc_ctx* ctx, int table_bytes, int hist_rows, int max_ops, aqc_loop** out);
int aqc_loop_destroy(aqc_ctx* ctx, aqc_loop* loop);
/* device address of the table: what aqc_reduce writes to (out_dev) and aqc_launch_ex reads from */
void* aqc_loop_table(aqc_loop* loop);
/* starts recording the body; `entry` (n_entry ops, must hold one AQS_SETCOND) is the program run
 * once before the loop: the `while` tool's condition */
int aqc_loop_begin(aqc_ctx* ctx, aqc_loop* loop, const struct aqs_op* entry, int n_entry);
/* while the loop records: queues a scalar program at this point of the body.  Otherwise the program
 * RUNS now, in stream order, on the table aqc_loop_start uploaded: the way a caller makes the first
 * pass of a loop tool by tool (the pass that builds neighbour lists and sizes scratch buffers)
 * without reading anything back, so that the device is still busy with it while the body is
 * recorded and instantiated. */
int aqc_loop_svm(aqc_ctx* ctx, aqc_loop* loop, const struct aqs_op* prog, int n);
/* clears the header and uploads the table (table_bytes of initial values, host) for programs run
 * directly and for an aqc_loop_run with table_in == NULL.  Not while recording. */
int aqc_loop_start(aqc_ctx* ctx, aqc_loop* loop, const void* table_in, uint32_t max_iters);
/* ends the recording and instantiates the graph.  The body must have queued a program with
 * AQS_SETCOND (a loop that cannot end is refused). */
int aqc_loop_end(aqc_ctx* ctx, aqc_loop* loop);
/* gives a recording up (also what aqc_loop_end does on failure) */
int aqc_loop_abort(aqc_ctx* ctx, aqc_loop* loop);
/* table_in: table_bytes of initial values (host), or NULL: header and table stay as aqc_loop_start
 * and the programs run since left them (max_iters is then the one given there).  When no body is
 * recorded (the recording failed) only the download happens.  max_iters bounds the loop (error
 * 0x30000 in the header when reached).  hdr_out, table_out (table_bytes), hist_out (hist_rows rows of
 * 16 + table_bytes: tool id, padding, snapshot) are host buffers, any may be NULL.  Syncs once. */
int aqc_loop_run(aqc_ctx* ctx, aqc_loop* loop, const void* table_in, uint32_t max_iters,
                 struct aqs_header* hdr_out, void* table_out, void* hist_out);
/* kernel nodes of the recorded body, and host milliseconds the last recording / instantiation took */
int aqc_loop_stats(const aqc_loop* loop, int* body_nodes, double* record_ms, double* instantiate_ms);
/* aqc_launch with scalars that live on the device: dev_scalars[k] != NULL makes the kernel read
 * scalar argument k from that device address when it RUNS instead of taking the value at args[k]
 * now (which is then only a fallback value and may be stale) -- what lets a recorded body read a
 * scalar that the loop itself updates (relax_midpoint of basic/time_scheme/midpoint.cl:141-157).
 * Only arguments flagged in aqc_kernel_dev_scalars(kernel_id) (bit k) may be bound that way. */
uint64_t aqc_kernel_dev_scalars(int kernel_id);
int aqc_launch_ex(aqc_ctx* ctx, int kernel_id, size_t n, void* const* args, int nargs,
                  const void* const* dev_scalars);

/* ---- measured FP32 (non-tensor) throughput of the device: two register-to-register FMA
 * micro-benchmarks (scalar FFMA and sm_100's packed FFMA2), TFLOP/s with fma = 2 flop.  The
 * denominator of the "% of FP32 peak" figures of the neighbour sweeps (no
 * such percentage without a measured peak); not a reference entry point. ---------------------- */
int aqc_fp32_peak(aqc_ctx* ctx, double* tflops_ffma, double* tflops_ffma2);

/* ---- events / profiling (Tool.cpp:296-310, Kernel.cpp:48-116) ------------ */
int aqc_event_create(aqc_ctx* ctx, void** ev);
int aqc_event_destroy(aqc_ctx* ctx, void* ev);
int aqc_event_record(aqc_ctx* ctx, void* ev);
int aqc_event_sync(aqc_ctx* ctx, void* ev);
int aqc_event_elapsed_ms(aqc_ctx* ctx, void* start, void* stop, float* ms);

#ifdef __cplusplus
}
#endif
#endif
