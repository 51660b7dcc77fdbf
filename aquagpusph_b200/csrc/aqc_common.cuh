// aqc_common.cuh -- internals shared by the translation units of libaquacuda.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "aquacuda.h"

struct aqc_ctx;
int aqc_fail(aqc_ctx* ctx, int code, const char* fmt, ...);

// Pair-mask cache of the v3 sweep engine (sweep.cuh, S3Cache): hit masks of the candidate filter,
// built once per geometry (positions + link-list + particle classes) and read by every sweep
// until one of those arrays is written through the library (aqc_pc_touch) or the caller says so
// (aqc_pairs_cache_invalidate).
constexpr size_t AQC_PC_ROUND_BYTES = 8 * 7 * 32 * 4; // one round of masks: tiles x consumer warps x lanes x 4 B
struct aqc_pair_cache {
    bool enabled = false;  // aqc_pairs_cache_enable
    bool valid = false;    // the masks belong to the key below
    bool unusable = false; // a CTA needed more passes than the pass table holds: sweeps filter
    const void *r = nullptr, *imove = nullptr, *icell = nullptr, *ihoc = nullptr; // key
    // the cache of the REMOTE (halo) sweeps: i cells from the local list, j rows / cells / heads from
    // the halo list (LINKLIST_REMOTE_PARAMS).  rj is part of the key but NOT of the write tracking:
    // cfd/MPI.cl::sort rewrites the halo positions every sub-iteration with the values they had,
    // and the halo list itself (icell, ihoc above) is only valid for the positions it was built from
    bool remote = false;
    const void *rj = nullptr, *icell_i = nullptr;
    uint32_t N = 0, nx = 0, ny = 0, nz = 0, nw = 0;
    int dims = 0;
    float cut2 = 0.f;
    uint32_t icls = 0, jcls = 0;           // particle classes (aqc_cls_bit) the masks were built for
    uint32_t icls_want = 0, jcls_want = 0; // union of the classes asked for so far
    uint32_t* masks = nullptr;
    size_t cap_rounds = 0;
    // neighbour lists (sweep.cuh, v4 engine): the default form of the cache; AQC_PAIR_LISTS=0 keeps
    // the hit masks of round 1 (MODE 1 / 2)
    bool lists = true;
    void* chunks = nullptr;    // uint2 [CTA][consumer warp][capc][32]
    size_t chunks_bytes = 0;
    uint8_t* cnt = nullptr;    // [round][consumer warp][32]
    uint32_t capc = 0;
    uint32_t* pass_tab = nullptr;
    size_t pass_cap = 0;
    unsigned long long* ctl = nullptr;      // device [4]
    unsigned long long* ctl_host = nullptr; // pinned [4]
    uint64_t builds = 0, hits = 0;
    // a build costs about half a sweep: pipelines whose geometry changes before a second sweep
    // reads the masks (served < 2, three times in a row) go without for a while
    uint32_t served = 0, poor_streak = 0, cooldown = 0;
};

// One slot per mpi-sync tool (aqc_mpi_sync_plan): what a call leaves behind so that the next call on
// a mask with the same content only gathers and exchanges (mpi.cu)
struct aqc_sync_plan {
    bool valid = false;
    const void* mask = nullptr;
    uint32_t n = 0;
    std::vector<const void*> fields;
    std::vector<size_t> elem_bytes;
    std::vector<char> peer;
    std::vector<uint32_t> soff, scnt, rcnt, roff; // per process
    uint32_t* perm = nullptr; // device: sort permutation of the mask
    size_t perm_cap = 0;
    uint32_t* mask_copy = nullptr; // device: the mask the plan was made from (AQC_MPI_VERIFY=1)
    size_t mask_copy_cap = 0;
    std::vector<std::pair<const char*, size_t>> deps; // ranges the mask is a function of
    uint64_t full = 0, reused = 0;
};

// Device ranges a caller wants to know about: dirty as soon as one of them is written through the
// library (aqc_watch_*: what lets a tool skip work whose inputs have not changed)
struct aqc_watch {
    std::vector<std::pair<const char*, size_t>> ranges;
    bool dirty = true;
};

struct aqc_ctx {
    aqc_pair_cache pc;
    aqc_pair_cache pcr; // remote (halo) sweeps
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;     // savers' downloads (aqc_side_*), created on first use
    cudaEvent_t side_fork_ev = nullptr;
    uint64_t launches = 0;
    aqc_defs defs{ 3, 1.f, 1.f, 1.f, 2.f, 3.f };
    // __DR_FACTOR__ / __MIN_BOUND_DIST__ / __ELASTIC_FACTOR__: script-specific defaults
    // (BIe/ElasticBounce.cl:31-45: 0.5, 0; Boundary/ElasticBounce.cl:31-58: 1.5, 0.3, 0)
    // unless the problem defines them (has_* set by aqc_set_define)
    float dr_factor = 0.5f;
    float min_bound_dist = 0.f;
    float elastic_factor = 0.f;
    bool lap_morris = false;        // __LAP_FORMULATION__ = __LAP_MORRIS__ (cfd/Interactions.cl:130-131)
    bool has_dr_factor = false, has_min_bound_dist = false;
    // TSCHEME_ADAMS_BASHFORTH_STEPS (adam_bashforth.cl:66-68: 5u unless the problem defines it)
    unsigned ab_steps = 5u;
    char err[512] = { 0 };

    // link-list scratch (grown on demand, never shrunk)
    uint32_t* sort_keys[2] = { nullptr, nullptr };
    uint32_t* sort_vals[2] = { nullptr, nullptr };
    size_t sort_cap = 0;
    uint32_t* sort_hist = nullptr; // digit totals + tile tickets + tile status words (linklist.cu)
    size_t sort_hist_cap = 0;
    bool sort_ghist_clean = false; // totals and tickets are zero (left so by the heads kernel)
    uint32_t* minmax_dev = nullptr; // 8 ordered-uint keys
    bool minmax_clean = false;      // ... hold the identities (left so by the prepare kernel)
    // per-cell mask of the particle classes present (sweep.cuh: aqc_cls_bit), rebuilt by the
    // sweeps whose j set is a small class (boundary elements): cells without one are skipped
    uint8_t* cell_cls = nullptr;
    size_t cell_cls_cap = 0;
    // packed j rows of the sweep being launched (sweep.cuh, s3_pack_kernel)
    void* pack_rows = nullptr;
    size_t pack_cap = 0;
    float* minmax_host = nullptr;   // pinned, 8 floats
    // reduction scratch
    void* red_dev = nullptr; // partials
    size_t red_cap = 0;
    void* red_host = nullptr; // pinned 64 B
    // multi-device (mpi.cu): one process per GPU, NCCL communicator loaded on demand
    int rank = 0, nranks = 1;
    void* comm = nullptr;              // ncclComm_t
    uint32_t* comm_counts = nullptr;   // device: [P] own counts + [P][P] gathered
    uint32_t* comm_counts_host = nullptr; // pinned mirror
    uint32_t* comm_perm = nullptr;     // sort permutation of the mask
    size_t comm_perm_cap = 0;
    void* comm_send = nullptr;         // packed outgoing fields
    size_t comm_send_cap = 0;
    bool comm_dead = false;            // the communicator was aborted (a peer is gone, a local fault)
    void* comm_watchdog = nullptr;     // backstop thread of the bounded waits (mpi.cu)
    std::vector<aqc_sync_plan> plans;  // aqc_mpi_sync_plan slots
    std::vector<aqc_watch> watches;    // aqc_watch_create slots
    // device-side loops (devloop.cu): the loop whose body is being recorded on `stream`, and the
    // device addresses of the scalar arguments of the launch in flight (aqc_launch_ex)
    struct aqc_loop* recording = nullptr;
    const void* const* dev_scalars = nullptr;
    // lanes (aqc_lane_*): `stream`, `cell_cls` and `pack_rows` above are those of the lane in use;
    // the other lane's are parked here (the sweeps' scratch is per lane: two sweeps may be in
    // flight at once)
    struct lane_state {
        cudaStream_t stream = nullptr;
        uint8_t* cell_cls = nullptr;
        size_t cell_cls_cap = 0;
        void* pack_rows = nullptr;
        size_t pack_cap = 0;
    } parked;
    int lane = 0;            // 0: the context's stream, 1: the branch stream
    bool lane1_made = false; // parked / current holds the branch stream
};
// a build of the pair cache rewrites what sweeps on the other lane may be reading: let them finish
static inline void aqc_lanes_drain_other(aqc_ctx* ctx)
{
    if (ctx->lane1_made && ctx->parked.stream && !ctx->recording)
        cudaStreamSynchronize(ctx->parked.stream);
}

int aqc_comm_minmax(aqc_ctx* ctx, uint32_t* keys); // mpi.cu
// Bounded wait for everything queued on the context's stream while a communicator is live: a
// collective whose peer died never completes, so the stream is polled against a deadline
// (AQC_COMM_TIMEOUT_S, default 60 s) together with ncclCommGetAsyncError; on expiry or error the
// communicator is aborted and the call fails (mpi.cu).  ev != nullptr: wait for that event only.
int aqc_comm_wait(aqc_ctx* ctx, cudaEvent_t ev);
void aqc_comm_abort(aqc_ctx* ctx); // mpi.cu
static inline int aqc_stream_wait(aqc_ctx* ctx)
{
    // a loop body is being recorded (devloop.cu): nothing runs, so there is nothing to wait for,
    // and whatever the caller wanted to read back is not there -- the recording fails cleanly
    if (ctx->recording)
        return aqc_fail(ctx, AQC_ERR_STATE, "this call synchronises with the device, which a recorded "
                                            "loop body cannot do");
    if (ctx->comm)
        return aqc_comm_wait(ctx, nullptr);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess)
        return aqc_fail(ctx, AQC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    return AQC_OK;
}
#define AQC_SYNC(ctx)                                                          \
    do {                                                                       \
        int s__ = aqc_stream_wait(ctx);                                        \
        if (s__)                                                               \
            return s__;                                                        \
    } while (0)

// [ptr, ptr + bytes) is about to be written: the pair-mask cache dies if it was built from it
static inline void aqc_pc_touch(aqc_ctx* ctx, const void* ptr, size_t bytes)
{
    aqc_pair_cache& c = ctx->pc;
    if (!ptr)
        return;
    const char* a = (const char*)ptr;
    const char* b = a + (bytes ? bytes : 1);
    for (aqc_watch& w : ctx->watches)
        if (!w.dirty)
            for (auto& d : w.ranges)
                if (a < d.first + d.second && d.first < b)
                    w.dirty = true;
    for (aqc_sync_plan& pl : ctx->plans) // mpi-sync plans die with the arrays their mask derives from
        if (pl.valid)
            for (auto& d : pl.deps)
                if (a < d.first + d.second && d.first < b)
                    pl.valid = false;
    auto hit = [&](const void* base, size_t n) {
        const char* x = (const char*)base;
        return base && a < x + n && x < b;
    };
    if (c.valid && (hit(c.r, (size_t)c.N * (c.dims == 3 ? 16 : 8)) || hit(c.imove, (size_t)c.N * 4) ||
                    hit(c.icell, (size_t)c.N * 4) || hit(c.ihoc, (size_t)c.nw * 4)))
        c.valid = false;
    aqc_pair_cache& q = ctx->pcr;
    if (q.valid && (hit(q.r, (size_t)q.N * (q.dims == 3 ? 16 : 8)) || hit(q.imove, (size_t)q.N * 4) ||
                    hit(q.icell_i, (size_t)q.N * 4) || hit(q.icell, (size_t)q.N * 4) ||
                    hit(q.ihoc, (size_t)q.nw * 4)))
        q.valid = false;
}
static inline void aqc_pc_invalidate(aqc_ctx* ctx) { ctx->pc.valid = ctx->pcr.valid = false; }
// bytes of one element of an array argument, from its reference type string ("vec*", "float*", ...)
static inline size_t aqc_type_bytes(const char* type, int dims)
{
    if (!strncmp(type, "matrix", 6))
        return dims == 3 ? 64 : 16;
    if (!strncmp(type, "vec4", 4) || !strncmp(type, "ivec4", 5) || !strncmp(type, "uivec4", 6) ||
        !strncmp(type, "svec4", 5))
        return 16;
    if (!strncmp(type, "vec2", 4) || !strncmp(type, "ivec2", 5) || !strncmp(type, "uivec2", 6))
        return 8;
    if (!strncmp(type, "vec", 3) || !strncmp(type, "ivec", 4) || !strncmp(type, "uivec", 5) ||
        !strncmp(type, "svec", 4))
        return dims == 3 ? 16 : 8;
    return 4; // float, int, uint, usize (32-bit indices)
}

#define AQC_CUDA(ctx, call)                                                    \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess)                                                \
            return aqc_fail((ctx), AQC_ERR_CUDA, "%s failed: %s (%s:%d)",      \
                            #call, cudaGetErrorString(e__), __FILE__,          \
                            __LINE__);                                         \
    } while (0)

#define AQC_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess)                                                \
            return aqc_fail((ctx), AQC_ERR_CUDA, "kernel launch failed: %s "   \
                            "(%s:%d)", cudaGetErrorString(e__), __FILE__,      \
                            __LINE__);                                         \
    } while (0)

static inline unsigned aqc_blocks(size_t n, unsigned bs)
{
    return (unsigned)((n + bs - 1) / bs);
}

// ---- kernel registry (registry.cu) ----------------------------------------
typedef int (*aqc_launcher)(aqc_ctx* ctx, size_t n, void* const* args);
struct aqc_kernel_entry {
    const char* script; // path relative to resources/Scripts/
    const char* entry;
    int dims; // 0 = both, 2, 3
    std::vector<aqc_arg_info> args;
    aqc_launcher fn;
    void* jit = nullptr; // run-time script (clc.cu): launched through aqc_script_launch, fn == nullptr
    uint64_t dev_mask = 0; // scalar arguments the launcher can bind to a device address (aqc_launch_ex)
};
int aqc_script_launch(aqc_ctx* ctx, const aqc_kernel_entry& e, size_t n, void* const* args); // clc.cu
std::vector<aqc_kernel_entry>& aqc_registry();
struct aqc_registrar {
    aqc_registrar(const char* script, const char* entry, int dims,
                  std::vector<aqc_arg_info> args, aqc_launcher fn, uint64_t dev_mask = 0)
    {
        aqc_registry().push_back({ script, entry, dims, std::move(args), fn, nullptr, dev_mask });
    }
};

// scalar argument helpers for launchers: args[k] is a host pointer
template <typename T>
static inline T aqc_scalar(void* const* args, int k)
{
    T v;
    memcpy(&v, args[k], sizeof(T));
    return v;
}
// ... or, under aqc_launch_ex, a device address the kernel reads when it runs (p != nullptr)
template <typename T> struct aqc_sv {
    const T* p;
    T v;
#if defined(__CUDACC__)
    __device__ __forceinline__ T get() const { return p ? *p : v; }
#endif
};
template <typename T>
static inline aqc_sv<T> aqc_scalar_sv(const aqc_ctx* ctx, void* const* args, int k)
{
    aqc_sv<T> s;
    s.p = ctx->dev_scalars ? (const T*)ctx->dev_scalars[k] : nullptr;
    s.v = aqc_scalar<T>(args, k);
    return s;
}
struct aqc_u4 { uint32_t x, y, z, w; };
struct aqc_f4 { float x, y, z, w; };
// vec scalar (g, domain_min, ...): 2 floats in 2-D, 4 in 3-D; returned padded
static inline aqc_f4 aqc_vec_scalar(void* const* args, int k, int dims)
{
    aqc_f4 v{ 0.f, 0.f, 0.f, 0.f };
    memcpy(&v, args[k], sizeof(float) * (dims == 3 ? 4 : 2));
    return v;
}
