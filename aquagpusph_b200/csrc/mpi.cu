// mpi.cu -- multi-device layer: the MPISync tool and the (added) all-reduces over
// NCCL on NVLink, one process per GPU.
//
// Replaces aquagpusph/CalcServer/MPISync.cpp:183-232 (sort the mask by
// destination process, gather the fields, per-process offset/count, exchange,
// mask set on receive; kernels MPISync.cl.in:31-80) and the host-staged MPI
// point-to-point of MPISync.cpp:564-638 / 932-1052 (blocking clEnqueueReadBuffer
// -> MPI_Isend / MPI_Recv -> clEnqueueWriteBuffer per field) with grouped
// ncclSend / ncclRecv directly between device buffers.
//
// NCCL is loaded with dlopen at aqc_comm_init, so libaquacuda.so itself has no
// link-time dependency on it and single-device runs never touch it.
#include <dlfcn.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <memory>
#include <thread>

#include "aqc_common.cuh"

namespace {

// ---- the few NCCL entry points used (ABI of nccl.h 2.x) ---------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };
enum { NCCL_CHAR = 0, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_FLOAT32 = 7 };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*CommAbort)(nccl_comm) = nullptr;
    int (*CommGetAsyncError)(nccl_comm, int*) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

const char* load_nccl()
{
    if (g_nccl.handle)
        return nullptr;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (auto n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
            break;
    if (!h)
        return "cannot dlopen libnccl.so.2";
#define SYM(field, name)                                                       \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                 \
    if (!g_nccl.field)                                                         \
        return "libnccl lacks " name;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(CommAbort, "ncclCommAbort")
    SYM(CommGetAsyncError, "ncclCommGetAsyncError")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return nullptr;
}

#define AQC_NCCL(ctx, call)                                                    \
    do {                                                                       \
        int r__ = (call);                                                      \
        if (r__ != 0)                                                          \
            return aqc_fail((ctx), AQC_ERR_NCCL, "%s failed: %s (%s:%d)",      \
                            #call, g_nccl.GetErrorString(r__), __FILE__,       \
                            __LINE__);                                         \
    } while (0)

// counts[p] = #{ i : mask[i] == p }  (Sender's n_send_mask + Reduction, MPISync.cpp:641-743)
__global__ void __launch_bounds__(256)
mask_count_kernel(const uint32_t* __restrict__ mask, uint32_t n, uint32_t nprocs,
                  uint32_t* __restrict__ counts)
{
    extern __shared__ uint32_t sc[];
    for (uint32_t k = threadIdx.x; k < nprocs; k += blockDim.x)
        sc[k] = 0;
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(mask + i);
        if (p < nprocs)
            atomicAdd(&sc[p], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < nprocs; k += blockDim.x)
        if (sc[k])
            atomicAdd(counts + k, sc[k]);
}

// dst[k] = src[perm[first + k]] for the sorted range [first, first + count): the
// UnSort of MPISync::setupFieldSort restricted to the elements that travel
template <typename T>
__global__ void __launch_bounds__(256)
gather_range_kernel(T* __restrict__ dst, const T* __restrict__ src,
                    const uint32_t* __restrict__ perm, uint32_t first, uint32_t count)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count)
        dst[k] = src[__ldg(perm + first + k)];
}

int gather_range(aqc_ctx* ctx, void* dst, const void* src, const uint32_t* perm, uint32_t first,
                 uint32_t count, size_t eb)
{
    if (!count)
        return AQC_OK;
    const unsigned g = aqc_blocks(count, 256);
    switch (eb) {
        case 4:
            gather_range_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((uint32_t*)dst, (const uint32_t*)src, perm, first, count);
            break;
        case 8:
            gather_range_kernel<uint2><<<g, 256, 0, ctx->stream>>>((uint2*)dst, (const uint2*)src, perm, first, count);
            break;
        case 16:
            gather_range_kernel<uint4><<<g, 256, 0, ctx->stream>>>((uint4*)dst, (const uint4*)src, perm, first, count);
            break;
        default:
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: unsupported element size %zu", eb);
    }
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

int grow(aqc_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (need <= *cap)
        return AQC_OK;
    if (*p) {
        AQC_SYNC(ctx);
        AQC_CUDA(ctx, cudaFree(*p));
    }
    *p = nullptr;
    *cap = 0;
    AQC_CUDA(ctx, cudaMalloc(p, need + need / 4 + 256));
    *cap = need + need / 4 + 256;
    return AQC_OK;
}

// flag |= 1 where a != b (AQC_MPI_VERIFY: a reused plan must see the mask it was made from)
__global__ void __launch_bounds__(256)
mask_diff_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t n,
                 uint32_t* __restrict__ flag)
{
    bool d = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        d |= a[i] != b[i];
    if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0)
        atomicOr(flag, 1u);
}

double now_s()
{
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

double comm_timeout_s();

// Backstop of the bounded waits.  aqc_comm_wait covers a host that waits for the device; a host
// that is stuck INSIDE an NCCL call (the enqueue of a collective whose peer process has died can
// block in the library), or that never comes back to a wait, is caught here: a thread per
// communicator watches the time since collective work was queued on a stream that has not
// drained.  After the time-out (+ a grace period that lets the in-line path report first) it
// says why on stderr, tries ncclCommAbort and ends the process with exit code 70 -- a job with a
// dead rank must end, not spin (round 1: three ranks held their GPUs for 870 s).
struct Watchdog {
    std::thread th;
    std::atomic<bool> stop{ false };
    std::atomic<long long> busy_since_ms{ 0 }; // 0: nothing pending
};

long long now_ms() { return (long long)(now_s() * 1e3); }

void watchdog_loop(aqc_ctx* ctx, Watchdog* w)
{
    cudaSetDevice(ctx->device);
    const long long limit_ms = (long long)((comm_timeout_s() + 5.0) * 1e3);
    while (!w->stop.load()) {
        usleep(200 * 1000);
        const long long b = w->busy_since_ms.load();
        if (!b || w->stop.load())
            continue;
        if (cudaStreamQuery(ctx->stream) == cudaSuccess) {
            long long expect = b;
            w->busy_since_ms.compare_exchange_strong(expect, 0);
            continue;
        }
        if (now_ms() - b > limit_ms) {
            fprintf(stderr, "aquacuda watchdog: rank %d of %d: collective work has been pending for %.0f s "
                            "(AQC_COMM_TIMEOUT_S + 5): a peer is gone or this rank is stuck inside NCCL; "
                            "aborting the communicator and ending the process (exit code 70)\n",
                    ctx->rank, ctx->nranks, (now_ms() - b) * 1e-3);
            fflush(stderr);
            void* c = ctx->comm;
            if (c && g_nccl.CommAbort)
                std::thread([c]() { g_nccl.CommAbort((nccl_comm)c); }).detach(); // (may block too)
            usleep(2000 * 1000);
            _exit(70);
        }
    }
}

void comm_mark_busy(aqc_ctx* ctx)
{
    Watchdog* w = (Watchdog*)ctx->comm_watchdog;
    if (!w)
        return;
    long long expect = 0;
    w->busy_since_ms.compare_exchange_strong(expect, now_ms());
}

void comm_mark_idle(aqc_ctx* ctx)
{
    Watchdog* w = (Watchdog*)ctx->comm_watchdog;
    if (w)
        w->busy_since_ms.store(0);
}

void watchdog_stop(aqc_ctx* ctx)
{
    Watchdog* w = (Watchdog*)ctx->comm_watchdog;
    if (!w)
        return;
    ctx->comm_watchdog = nullptr;
    w->stop.store(true);
    if (w->th.joinable() && w->th.get_id() != std::this_thread::get_id())
        w->th.join();
    else if (w->th.joinable())
        w->th.detach();
    delete w;
}

double comm_timeout_s()
{
    static double v = -1.0;
    if (v < 0.0) {
        const char* e = getenv("AQC_COMM_TIMEOUT_S");
        v = (e && atof(e) > 0.0) ? atof(e) : 60.0;
    }
    return v;
}

bool plans_enabled() // AQC_MPI_PLANS=0: every mpi-sync call sorts and counts (A/B runs)
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AQC_MPI_PLANS");
        v = (e && !atoi(e)) ? 0 : 1;
    }
    return v == 1;
}

bool verify_plans()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AQC_MPI_VERIFY");
        v = (e && atoi(e)) ? 1 : 0;
    }
    return v == 1;
}

constexpr size_t SEND_ALIGN = 256; // every field block of the packed send buffer starts on this
size_t align_up(size_t x) { return (x + SEND_ALIGN - 1) & ~(SEND_ALIGN - 1); }

} // namespace

// A peer that died (or faulted) never completes its side of a collective: without this the
// survivors spin inside ncclRecv for ever -- NCCL is dlopen'ed here, nobody else watches it.
void aqc_comm_abort(aqc_ctx* ctx)
{
    if (!ctx || !ctx->comm)
        return;
    nccl_comm c = (nccl_comm)ctx->comm;
    ctx->comm = nullptr;
    ctx->comm_dead = true;
    comm_mark_idle(ctx); // (the watchdog stands down: the failure is being reported in line)
    // ncclCommAbort also ends the kernels of this rank that wait for the peer; it can block itself
    // when the peer process is gone, so it gets five seconds on a thread of its own
    auto done = std::make_shared<std::atomic<bool>>(false);
    std::thread([c, done]() {
        g_nccl.CommAbort(c);
        done->store(true);
    }).detach();
    for (int k = 0; k < 500 && !done->load(); k++)
        usleep(10 * 1000);
}

int aqc_comm_wait(aqc_ctx* ctx, cudaEvent_t ev)
{
    const double t0 = now_s(), limit = comm_timeout_s();
    unsigned polls = 0;
    for (;;) {
        const cudaError_t e = ev ? cudaEventQuery(ev) : cudaStreamQuery(ctx->stream);
        if (e == cudaSuccess) {
            if (!ev)
                comm_mark_idle(ctx);
            return AQC_OK;
        }
        if (e != cudaErrorNotReady) {
            char msg[256];
            snprintf(msg, sizeof(msg), "%s", cudaGetErrorString(e));
            aqc_comm_abort(ctx);
            return aqc_fail(ctx, AQC_ERR_CUDA, "rank %d: device fault while a collective was pending: %s",
                            ctx->rank, msg);
        }
        if (ctx->comm && !(++polls & 1023u)) {
            int st = 0;
            // ncclInProgress (7) is what a non-blocking communicator reports while it works
            if (g_nccl.CommGetAsyncError((nccl_comm)ctx->comm, &st) == 0 && st != 0 && st != 7) {
                aqc_comm_abort(ctx);
                return aqc_fail(ctx, AQC_ERR_NCCL, "rank %d: NCCL reports an asynchronous error: %s",
                                ctx->rank, g_nccl.GetErrorString(st));
            }
            if (now_s() - t0 > limit) {
                aqc_comm_abort(ctx);
                return aqc_fail(ctx, AQC_ERR_NCCL,
                                "rank %d: nothing completed on the stream for %.0f s with a collective "
                                "pending: a peer is gone; communicator aborted (AQC_COMM_TIMEOUT_S)",
                                ctx->rank, limit);
            }
        }
        if (polls > 20000u) { // a long wait: stop burning the core
            timespec ts{ 0, 50000 };
            nanosleep(&ts, nullptr);
        }
    }
}

extern "C" int aqc_comm_unique_id(void* id_out)
{
    if (!id_out)
        return AQC_ERR_ARG;
    if (load_nccl())
        return AQC_ERR_NCCL;
    nccl_uid id;
    if (g_nccl.GetUniqueId(&id))
        return AQC_ERR_NCCL;
    memcpy(id_out, &id, sizeof(id));
    return AQC_OK;
}

extern "C" int aqc_comm_init(aqc_ctx* ctx, int rank, int size, const void* unique_id)
{
    if (!ctx || rank < 0 || size < 1 || rank >= size)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_comm_init: bad rank/size %d/%d", rank, size);
    ctx->rank = rank;
    ctx->nranks = size;
    ctx->comm_dead = false;
    if (size == 1)
        return AQC_OK; // nothing to talk to
    if (!unique_id)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_comm_init: NULL unique id");
    if (const char* why = load_nccl())
        return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_comm_init: %s (%s)", why, dlerror());
    AQC_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, unique_id, sizeof(id));
    nccl_comm comm = nullptr;
    AQC_NCCL(ctx, g_nccl.CommInitRank(&comm, size, id, rank));
    ctx->comm = comm;
    {
        const char* e = getenv("AQC_COMM_WATCHDOG");
        if (!(e && atoi(e) == 0)) {
            Watchdog* w = new Watchdog();
            ctx->comm_watchdog = w;
            w->th = std::thread(watchdog_loop, ctx, w);
        }
    }
    // [2P] own counts + peer flags, then [P][2P] gathered
    AQC_CUDA(ctx, cudaMalloc(&ctx->comm_counts, (size_t)2 * size * (size + 1) * sizeof(uint32_t) + 256));
    AQC_CUDA(ctx, cudaMallocHost(&ctx->comm_counts_host, (size_t)2 * size * (size + 1) * sizeof(uint32_t)));
    return AQC_OK;
}

extern "C" int aqc_comm_destroy(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    watchdog_stop(ctx);
    if (ctx->comm) {
        // a clean shutdown drains the stream first; if that does not end (a peer left
        // without its matching call) the communicator is aborted instead of destroyed
        if (aqc_comm_wait(ctx, nullptr) == AQC_OK && ctx->comm)
            g_nccl.CommDestroy((nccl_comm)ctx->comm);
        ctx->comm = nullptr;
    }
    cudaFree(ctx->comm_counts);
    cudaFreeHost(ctx->comm_counts_host);
    cudaFree(ctx->comm_perm);
    cudaFree(ctx->comm_send);
    for (aqc_sync_plan& pl : ctx->plans) {
        cudaFree(pl.perm);
        cudaFree(pl.mask_copy);
        pl = aqc_sync_plan();
    }
    ctx->comm_counts = nullptr;
    ctx->comm_counts_host = nullptr;
    ctx->comm_perm = nullptr;
    ctx->comm_send = nullptr;
    ctx->comm_perm_cap = ctx->comm_send_cap = 0;
    return AQC_OK;
}

extern "C" int aqc_comm_rank(const aqc_ctx* ctx) { return ctx ? ctx->rank : 0; }
extern "C" int aqc_comm_size(const aqc_ctx* ctx) { return ctx ? ctx->nranks : 1; }

extern "C" int aqc_mpi_sync_plan(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    ctx->plans.emplace_back();
    return (int)ctx->plans.size() - 1;
}

extern "C" int aqc_mpi_sync_stats(const aqc_ctx* ctx, int plan, uint64_t* full, uint64_t* reused)
{
    if (!ctx || plan < 0 || plan >= (int)ctx->plans.size())
        return AQC_ERR_ARG;
    if (full)
        *full = ctx->plans[plan].full;
    if (reused)
        *reused = ctx->plans[plan].reused;
    return AQC_OK;
}

namespace {

// Steps 3-5 of MPISync::_execute for a known layout: pack what travels, exchange, rewrite the
// mask.  Nothing here waits for the device.
int sync_exchange(aqc_ctx* ctx, const aqc_sync_plan& pl, const uint32_t* perm, aqc_usize* mask,
                  aqc_usize n, int nfields, void* const* fields, const size_t* elem_bytes,
                  aqc_usize* n_received)
{
    const int P = ctx->nranks, me = ctx->rank;
    size_t total_send = 0;
    std::vector<uint32_t> spack(P + 1, 0); // position of p's block inside a field's send block
    for (int p = 0; p < P; p++) {
        spack[p + 1] = spack[p] + (pl.peer[p] ? pl.scnt[p] : 0);
        if (pl.peer[p])
            total_send += pl.scnt[p];
    }
    // one block per field, each starting on a 256-byte boundary: the 4-byte mpi_iset block is
    // followed by 16-byte vectors, and the gather stores whole elements (uint2 / uint4)
    std::vector<size_t> fbase(nfields + 1, 0);
    for (int f = 0; f < nfields; f++)
        fbase[f + 1] = fbase[f] + align_up(total_send * elem_bytes[f]);
    int rc = grow(ctx, &ctx->comm_send, &ctx->comm_send_cap, fbase[nfields] + SEND_ALIGN);
    if (rc)
        return rc;
    for (int f = 0; f < nfields; f++)
        for (int p = 0; p < P; p++)
            if (pl.peer[p] && pl.scnt[p]) {
                char* dst = (char*)ctx->comm_send + fbase[f] + (size_t)spack[p] * elem_bytes[f];
                rc = gather_range(ctx, dst, fields[f], perm, pl.soff[p], pl.scnt[p], elem_bytes[f]);
                if (rc)
                    return rc;
            }
    // exchange, device to device.  A failure between GroupStart and GroupEnd still closes the
    // group, and any NCCL failure aborts the communicator: the peers' waits then end too.
    comm_mark_busy(ctx);
    int nrc = g_nccl.GroupStart();
    bool opened = nrc == 0;
    for (int p = 0; p < P && nrc == 0; p++) {
        if (!pl.peer[p])
            continue;
        for (int f = 0; f < nfields && nrc == 0; f++) {
            if (pl.scnt[p]) {
                const char* src = (const char*)ctx->comm_send + fbase[f] + (size_t)spack[p] * elem_bytes[f];
                nrc = g_nccl.Send(src, (size_t)pl.scnt[p] * elem_bytes[f], NCCL_CHAR, p,
                                  (nccl_comm)ctx->comm, ctx->stream);
            }
            if (nrc == 0 && pl.rcnt[p]) {
                char* dst = (char*)fields[f] + (size_t)pl.roff[p] * elem_bytes[f];
                nrc = g_nccl.Recv(dst, (size_t)pl.rcnt[p] * elem_bytes[f], NCCL_CHAR, p,
                                  (nccl_comm)ctx->comm, ctx->stream);
            }
        }
    }
    if (opened) {
        const int erc = g_nccl.GroupEnd();
        if (nrc == 0)
            nrc = erc;
    }
    if (nrc != 0) {
        char msg[160];
        snprintf(msg, sizeof(msg), "%s", g_nccl.GetErrorString(nrc));
        aqc_comm_abort(ctx);
        return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_mpi_sync: rank %d: exchange failed: %s", me, msg);
    }
    // mask = own rank everywhere, then the sender's rank over every received block
    // (MPISync.cpp:222-223 + set_mask, MPISync.cl.in:68-80)
    const uint32_t mine = (uint32_t)me;
    rc = aqc_fill(ctx, mask, n, sizeof(uint32_t), &mine);
    if (rc)
        return rc;
    for (int p = 0; p < P; p++)
        if (pl.rcnt[p]) {
            const uint32_t v = (uint32_t)p;
            rc = aqc_fill(ctx, mask + pl.roff[p], pl.rcnt[p], sizeof(uint32_t), &v);
            if (rc)
                return rc;
        }
    if (n_received)
        *n_received = pl.roff[P];
    return AQC_OK;
}

} // namespace

extern "C" int aqc_mpi_sync_ex(aqc_ctx* ctx, int plan, aqc_usize* mask, aqc_usize n, int nfields,
                               void* const* fields, const size_t* elem_bytes, int nprocs,
                               const unsigned* procs, aqc_usize* n_received, int ndeps,
                               const void* const* dep_ptrs, const size_t* dep_bytes)
{
    if (n_received)
        *n_received = 0;
    if (!ctx || !mask || nfields < 0 || (nfields && (!fields || !elem_bytes)) ||
        (ndeps > 0 && (!dep_ptrs || !dep_bytes)))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: NULL argument");
    if (plan >= (int)ctx->plans.size())
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: unknown plan %d", plan);
    if (ctx->comm_dead)
        return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_mpi_sync: rank %d: the communicator was aborted", ctx->rank);
    // MPISync.cpp:186-187: nobody to talk to => nothing happens (the mask is left alone)
    if (ctx->nranks <= 1 || !ctx->comm || !n)
        return AQC_OK;
    for (int f = 0; f < nfields; f++)
        if (elem_bytes[f] != 4 && elem_bytes[f] != 8 && elem_bytes[f] != 16)
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: unsupported element size %zu", elem_bytes[f]);
    const int P = ctx->nranks, me = ctx->rank;
    std::vector<char> peer(P, procs ? 0 : 1);
    if (procs)
        for (int k = 0; k < nprocs; k++) {
            if (procs[k] >= (unsigned)P)
                return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: process %u out of range", procs[k]);
            peer[procs[k]] = 1;
        }
    peer[me] = 0;

    aqc_sync_plan scratch;
    aqc_sync_plan* pl = plan >= 0 ? &ctx->plans[plan] : &scratch;
    // ---- the plan of the previous call stands when the caller declared what the mask derives
    // from, none of that was written since, and the call is the same one
    bool reuse = plan >= 0 && plans_enabled() && pl->valid && pl->mask == mask && pl->n == n && pl->peer == peer &&
                 (int)pl->fields.size() == nfields;
    for (int f = 0; reuse && f < nfields; f++)
        reuse = pl->fields[f] == fields[f] && pl->elem_bytes[f] == elem_bytes[f];
    if (reuse && verify_plans()) {
        AQC_CUDA(ctx, cudaMemsetAsync(ctx->comm_counts, 0, sizeof(uint32_t), ctx->stream));
        const unsigned cap = (unsigned)ctx->sm_count * 4, g = aqc_blocks(n, 256 * 8);
        mask_diff_kernel<<<g > cap ? cap : g, 256, 0, ctx->stream>>>(mask, pl->mask_copy, n, ctx->comm_counts);
        AQC_LAUNCH_CHECK(ctx);
        AQC_CUDA(ctx, cudaMemcpyAsync(ctx->comm_counts_host, ctx->comm_counts, sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        AQC_SYNC(ctx);
        if (ctx->comm_counts_host[0])
            return aqc_fail(ctx, AQC_ERR_STATE, "aqc_mpi_sync: rank %d: plan %d reused on a mask that "
                            "changed (the tool's depends list is incomplete)", me, plan);
    }
    bool was_valid = pl->valid;
    pl->valid = false; // (the touches below hit the plan's own ranges when a field is a dependency)
    aqc_pc_touch(ctx, mask, (size_t)n * sizeof(aqc_usize)); // the mask and every field are rewritten
    for (int f = 0; f < nfields; f++)
        aqc_pc_touch(ctx, fields[f], (size_t)n * elem_bytes[f]);
    if (reuse) {
        pl->valid = was_valid;
        pl->reused++;
        return sync_exchange(ctx, *pl, pl->perm, mask, n, nfields, fields, elem_bytes, n_received);
    }

    // 1. stable sort of the mask by destination (RadixSort of MPISync::setupSort);
    //    perm[k] = original index of the k-th sorted element
    uint32_t** permp = plan >= 0 ? &pl->perm : &ctx->comm_perm;
    size_t* permcap = plan >= 0 ? &pl->perm_cap : &ctx->comm_perm_cap;
    int rc = grow(ctx, (void**)permp, permcap, (size_t)n * sizeof(uint32_t));
    if (rc)
        return rc;
    if (plan >= 0 && verify_plans()) {
        rc = grow(ctx, (void**)&pl->mask_copy, &pl->mask_copy_cap, (size_t)n * sizeof(uint32_t));
        if (rc)
            return rc;
        AQC_CUDA(ctx, cudaMemcpyAsync(pl->mask_copy, mask, (size_t)n * sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    }
    rc = aqc_radix_sort(ctx, mask, n, (aqc_usize)P, *permp, nullptr);
    if (rc)
        return rc;
    // 2. per-destination counts and the peer list, shared with every rank (one all-gather
    //    instead of P-1 tagged count messages, MPISync.cpp:581,948)
    uint32_t* d_mine = ctx->comm_counts;         // [P] counts, [P] peer flags
    uint32_t* d_all = ctx->comm_counts + 2 * P;  // [P][2P]
    AQC_CUDA(ctx, cudaMemsetAsync(d_mine, 0, P * sizeof(uint32_t), ctx->stream));
    {
        std::vector<uint32_t>& stage = pl->soff; // (scratch: rewritten below)
        stage.assign(P, 0);
        for (int p = 0; p < P; p++)
            stage[p] = (uint32_t)peer[p];
        memcpy(ctx->comm_counts_host, stage.data(), P * sizeof(uint32_t));
        AQC_CUDA(ctx, cudaMemcpyAsync(d_mine + P, ctx->comm_counts_host, P * sizeof(uint32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
        unsigned g = aqc_blocks(n, 256 * 8);
        const unsigned cap = (unsigned)ctx->sm_count * 4;
        mask_count_kernel<<<g > cap ? cap : g, 256, P * sizeof(uint32_t), ctx->stream>>>(mask, n, P, d_mine);
        AQC_LAUNCH_CHECK(ctx);
    }
    {
        comm_mark_busy(ctx);
        const int nrc = g_nccl.AllGather(d_mine, d_all, 2 * P, NCCL_UINT32, (nccl_comm)ctx->comm, ctx->stream);
        if (nrc != 0) {
            aqc_comm_abort(ctx);
            return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_mpi_sync: rank %d: count all-gather failed: %s", me,
                            g_nccl.GetErrorString(nrc));
        }
    }
    // (the pinned mirror is read by the copy above before this one overwrites it: stream order)
    AQC_CUDA(ctx, cudaMemcpyAsync(ctx->comm_counts_host, d_all, (size_t)2 * P * P * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    AQC_SYNC(ctx);
    const uint32_t* all = ctx->comm_counts_host; // all[src * 2P + dst] counts, all[src * 2P + P + dst] flags
    pl->soff.assign(P + 1, 0);
    pl->scnt.assign(P, 0);
    pl->rcnt.assign(P, 0);
    pl->roff.assign(P + 1, 0);
    for (int p = 0; p < P; p++) {
        // both sides must agree that they talk to each other, or one of them would wait for a
        // message that is never sent (the reference's `processes` attribute is per tool and rank)
        if (p != me && (all[p * 2 * P + P + me] != 0) != (peer[p] != 0))
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: rank %d %s rank %d, which %s it: the process "
                            "lists must be symmetric", me, peer[p] ? "lists" : "does not list", p,
                            peer[p] ? "does not list" : "lists");
        pl->scnt[p] = all[me * 2 * P + p];
        pl->soff[p + 1] = pl->soff[p] + pl->scnt[p]; // sorted position of the first element bound to p
        pl->rcnt[p] = peer[p] ? all[p * 2 * P + me] : 0;
    }
    for (int p = 0; p < P; p++)
        pl->roff[p + 1] = pl->roff[p] + pl->rcnt[p]; // received blocks are packed from 0 in process order
    if (pl->roff[P] > n)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_mpi_sync: %u elements received but the arrays hold %u",
                        pl->roff[P], n);
    pl->peer = peer;
    rc = sync_exchange(ctx, *pl, *permp, mask, n, nfields, fields, elem_bytes, n_received);
    if (rc)
        return rc;
    pl->full++;
    if (plan >= 0 && ndeps > 0) {
        pl->mask = mask;
        pl->n = n;
        pl->fields.assign(fields, fields + nfields);
        pl->elem_bytes.assign(elem_bytes, elem_bytes + nfields);
        pl->deps.clear();
        for (int k = 0; k < ndeps; k++)
            if (dep_ptrs[k])
                pl->deps.emplace_back((const char*)dep_ptrs[k], dep_bytes[k]);
        pl->valid = true;
    }
    return AQC_OK;
}

extern "C" int aqc_mpi_sync(aqc_ctx* ctx, aqc_usize* mask, aqc_usize n, int nfields,
                            void* const* fields, const size_t* elem_bytes, int nprocs,
                            const unsigned* procs, aqc_usize* n_received)
{
    return aqc_mpi_sync_ex(ctx, -1, mask, n, nfields, fields, elem_bytes, nprocs, procs, n_received, 0,
                           nullptr, nullptr);
}

extern "C" int aqc_allreduce(aqc_ctx* ctx, int op, int type, void* dev_inout, size_t count)
{
    if (!ctx || !dev_inout)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_allreduce: NULL argument");
    if (ctx->comm_dead)
        return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_allreduce: rank %d: the communicator was aborted", ctx->rank);
    if (ctx->nranks <= 1 || !ctx->comm || !count)
        return AQC_OK;
    aqc_pc_touch(ctx, dev_inout, count * (type == AQC_T_VEC4 ? 16 : type == AQC_T_VEC2 ? 8 : 4));
    int nt, ncomp = 1;
    switch (type) {
        case AQC_T_F32: nt = NCCL_FLOAT32; break;
        case AQC_T_U32: nt = NCCL_UINT32; break;
        case AQC_T_I32: nt = NCCL_INT32; break;
        case AQC_T_VEC2: nt = NCCL_FLOAT32; ncomp = 2; break;
        case AQC_T_VEC4: nt = NCCL_FLOAT32; ncomp = 4; break;
        default: return aqc_fail(ctx, AQC_ERR_ARG, "aqc_allreduce: unknown type %d", type);
    }
    const int nop = op == AQC_OP_SUM ? NCCL_SUM : (op == AQC_OP_MIN ? NCCL_MIN : NCCL_MAX);
    comm_mark_busy(ctx);
    const int nrc = g_nccl.AllReduce(dev_inout, dev_inout, count * ncomp, nt, nop, (nccl_comm)ctx->comm,
                                     ctx->stream);
    if (nrc != 0) {
        aqc_comm_abort(ctx);
        return aqc_fail(ctx, AQC_ERR_NCCL, "aqc_allreduce: rank %d: %s", ctx->rank, g_nccl.GetErrorString(nrc));
    }
    return AQC_OK;
}

extern "C" int aqc_allreduce_host(aqc_ctx* ctx, int op, int type, void* host_inout, size_t count)
{
    if (!ctx || !host_inout)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_allreduce_host: NULL argument");
    if (ctx->nranks <= 1 || !ctx->comm || !count)
        return AQC_OK;
    const size_t eb = (type == AQC_T_VEC2 ? 8 : (type == AQC_T_VEC4 ? 16 : 4));
    if (count * eb > 64)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_allreduce_host: at most 64 bytes");
    // the context's 8-word min/max scratch doubles as the staging buffer (64 B pinned + 32 B device
    // is too small for 64 B: use the count table)
    void* d = ctx->comm_counts;
    AQC_CUDA(ctx, cudaMemcpyAsync(d, host_inout, count * eb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = aqc_allreduce(ctx, op, type, d, count);
    if (rc)
        return rc;
    AQC_CUDA(ctx, cudaMemcpyAsync(ctx->red_host, d, count * eb, cudaMemcpyDeviceToHost, ctx->stream));
    AQC_SYNC(ctx);
    memcpy(host_inout, ctx->red_host, count * eb);
    return AQC_OK;
}

// Used by aqc_linklist_build: every rank hashes on ONE global grid, so that cell
// indices are comparable with the single-device run and halo particles never fall
// below the local r_min (the reference converts a negative float to unsigned there,
// LinkList.cl.in:74-77).  keys = 4 ordered-uint minima followed by 4 maxima.
int aqc_comm_minmax(aqc_ctx* ctx, uint32_t* keys)
{
    if (ctx->comm_dead)
        return aqc_fail(ctx, AQC_ERR_NCCL, "link-list: rank %d: the communicator was aborted", ctx->rank);
    if (ctx->nranks <= 1 || !ctx->comm)
        return AQC_OK;
    comm_mark_busy(ctx);
    int nrc = g_nccl.GroupStart();
    if (nrc == 0) {
        nrc = g_nccl.AllReduce(keys, keys, 4, NCCL_UINT32, NCCL_MIN, (nccl_comm)ctx->comm, ctx->stream);
        if (nrc == 0)
            nrc = g_nccl.AllReduce(keys + 4, keys + 4, 4, NCCL_UINT32, NCCL_MAX, (nccl_comm)ctx->comm,
                                   ctx->stream);
        const int erc = g_nccl.GroupEnd();
        if (nrc == 0)
            nrc = erc;
    }
    if (nrc != 0) {
        aqc_comm_abort(ctx);
        return aqc_fail(ctx, AQC_ERR_NCCL, "link-list: rank %d: min/max all-reduce failed: %s", ctx->rank,
                        g_nccl.GetErrorString(nrc));
    }
    return AQC_OK;
}
