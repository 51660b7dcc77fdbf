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
        AQC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        AQC_CUDA(ctx, cudaFree(*p));
    }
    *p = nullptr;
    *cap = 0;
    AQC_CUDA(ctx, cudaMalloc(p, need + need / 4 + 256));
    *cap = need + need / 4 + 256;
    return AQC_OK;
}

} // namespace

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
    AQC_CUDA(ctx, cudaMalloc(&ctx->comm_counts, (size_t)size * (size + 2) * sizeof(uint32_t) + 256));
    AQC_CUDA(ctx, cudaMallocHost(&ctx->comm_counts_host, (size_t)size * (size + 2) * sizeof(uint32_t)));
    return AQC_OK;
}

extern "C" int aqc_comm_destroy(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (ctx->comm) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((nccl_comm)ctx->comm);
        ctx->comm = nullptr;
    }
    cudaFree(ctx->comm_counts);
    cudaFreeHost(ctx->comm_counts_host);
    cudaFree(ctx->comm_perm);
    cudaFree(ctx->comm_send);
    ctx->comm_counts = nullptr;
    ctx->comm_counts_host = nullptr;
    ctx->comm_perm = nullptr;
    ctx->comm_send = nullptr;
    ctx->comm_perm_cap = ctx->comm_send_cap = 0;
    return AQC_OK;
}

extern "C" int aqc_comm_rank(const aqc_ctx* ctx) { return ctx ? ctx->rank : 0; }
extern "C" int aqc_comm_size(const aqc_ctx* ctx) { return ctx ? ctx->nranks : 1; }

extern "C" int aqc_mpi_sync(aqc_ctx* ctx, aqc_usize* mask, aqc_usize n, int nfields,
                            void* const* fields, const size_t* elem_bytes, int nprocs,
                            const unsigned* procs, aqc_usize* n_received)
{
    if (n_received)
        *n_received = 0;
    if (!ctx || !mask || nfields < 0 || (nfields && (!fields || !elem_bytes)))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: NULL argument");
    // MPISync.cpp:186-187: nobody to talk to => nothing happens (the mask is left alone)
    if (ctx->nranks <= 1 || !ctx->comm || !n)
        return AQC_OK;
    aqc_pc_touch(ctx, mask, (size_t)n * sizeof(aqc_usize)); // the mask and every field are rewritten
    for (int f = 0; f < nfields; f++)
        aqc_pc_touch(ctx, fields[f], (size_t)n * elem_bytes[f]);
    const int P = ctx->nranks, me = ctx->rank;
    std::vector<char> peer(P, procs ? 0 : 1);
    if (procs)
        for (int k = 0; k < nprocs; k++) {
            if (procs[k] >= (unsigned)P)
                return aqc_fail(ctx, AQC_ERR_ARG, "aqc_mpi_sync: process %u out of range", procs[k]);
            peer[procs[k]] = 1;
        }
    peer[me] = 0;

    // 1. stable sort of the mask by destination (RadixSort of MPISync::setupSort);
    //    perm[k] = original index of the k-th sorted element
    int rc = grow(ctx, (void**)&ctx->comm_perm, &ctx->comm_perm_cap, (size_t)n * sizeof(uint32_t));
    if (rc)
        return rc;
    rc = aqc_radix_sort(ctx, mask, n, (aqc_usize)P, ctx->comm_perm, nullptr);
    if (rc)
        return rc;
    // 2. per-destination counts, shared with every rank (one all-gather instead of
    //    P-1 tagged count messages, MPISync.cpp:581,948)
    uint32_t* d_mine = ctx->comm_counts;              // [P]
    uint32_t* d_all = ctx->comm_counts + P;           // [P][P]
    AQC_CUDA(ctx, cudaMemsetAsync(d_mine, 0, P * sizeof(uint32_t), ctx->stream));
    {
        unsigned g = aqc_blocks(n, 256 * 8);
        const unsigned cap = (unsigned)ctx->sm_count * 4;
        mask_count_kernel<<<g > cap ? cap : g, 256, P * sizeof(uint32_t), ctx->stream>>>(mask, n, P, d_mine);
        AQC_LAUNCH_CHECK(ctx);
    }
    AQC_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, P, NCCL_UINT32, (nccl_comm)ctx->comm, ctx->stream));
    AQC_CUDA(ctx, cudaMemcpyAsync(ctx->comm_counts_host, d_all, (size_t)P * P * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    AQC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t* all = ctx->comm_counts_host; // all[src * P + dst]
    std::vector<uint32_t> soff(P + 1, 0), scnt(P, 0), rcnt(P, 0), roff(P + 1, 0);
    for (int p = 0; p < P; p++) {
        scnt[p] = all[me * P + p];
        soff[p + 1] = soff[p] + scnt[p]; // sorted position of the first element bound to p
        rcnt[p] = peer[p] ? all[p * P + me] : 0;
    }
    for (int p = 0; p < P; p++)
        roff[p + 1] = roff[p] + rcnt[p]; // received blocks are packed from 0 in process order
    if (roff[P] > n)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_mpi_sync: %u elements received but the arrays hold %u",
                        roff[P], n);
    // 3. pack what travels (fields gathered in sorted order, only the ranges bound to peers)
    size_t total_send = 0;
    for (int p = 0; p < P; p++)
        if (peer[p])
            total_send += scnt[p];
    size_t bytes_per_elem = 0;
    for (int f = 0; f < nfields; f++)
        bytes_per_elem += elem_bytes[f];
    rc = grow(ctx, &ctx->comm_send, &ctx->comm_send_cap, total_send * bytes_per_elem + 16);
    if (rc)
        return rc;
    std::vector<size_t> fbase(nfields + 1, 0);
    for (int f = 0; f < nfields; f++)
        fbase[f + 1] = fbase[f] + total_send * elem_bytes[f];
    std::vector<uint32_t> spack(P + 1, 0); // position of p's block inside a field's send buffer
    for (int p = 0; p < P; p++)
        spack[p + 1] = spack[p] + (peer[p] ? scnt[p] : 0);
    for (int f = 0; f < nfields; f++)
        for (int p = 0; p < P; p++)
            if (peer[p] && scnt[p]) {
                char* dst = (char*)ctx->comm_send + fbase[f] + (size_t)spack[p] * elem_bytes[f];
                rc = gather_range(ctx, dst, fields[f], ctx->comm_perm, soff[p], scnt[p], elem_bytes[f]);
                if (rc)
                    return rc;
            }
    // 4. exchange, device to device
    AQC_NCCL(ctx, g_nccl.GroupStart());
    for (int p = 0; p < P; p++) {
        if (!peer[p])
            continue;
        for (int f = 0; f < nfields; f++) {
            if (scnt[p]) {
                const char* src = (const char*)ctx->comm_send + fbase[f] + (size_t)spack[p] * elem_bytes[f];
                AQC_NCCL(ctx, g_nccl.Send(src, (size_t)scnt[p] * elem_bytes[f], NCCL_CHAR, p,
                                          (nccl_comm)ctx->comm, ctx->stream));
            }
            if (rcnt[p]) {
                char* dst = (char*)fields[f] + (size_t)roff[p] * elem_bytes[f];
                AQC_NCCL(ctx, g_nccl.Recv(dst, (size_t)rcnt[p] * elem_bytes[f], NCCL_CHAR, p,
                                          (nccl_comm)ctx->comm, ctx->stream));
            }
        }
    }
    AQC_NCCL(ctx, g_nccl.GroupEnd());
    // 5. mask = own rank everywhere, then the sender's rank over every received block
    //    (MPISync.cpp:222-223 + set_mask, MPISync.cl.in:68-80)
    const uint32_t mine = (uint32_t)me;
    rc = aqc_fill(ctx, mask, n, sizeof(uint32_t), &mine);
    if (rc)
        return rc;
    for (int p = 0; p < P; p++)
        if (rcnt[p]) {
            const uint32_t v = (uint32_t)p;
            rc = aqc_fill(ctx, mask + roff[p], rcnt[p], sizeof(uint32_t), &v);
            if (rc)
                return rc;
        }
    if (n_received)
        *n_received = roff[P];
    return AQC_OK;
}

extern "C" int aqc_allreduce(aqc_ctx* ctx, int op, int type, void* dev_inout, size_t count)
{
    if (!ctx || !dev_inout)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_allreduce: NULL argument");
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
    AQC_NCCL(ctx, g_nccl.AllReduce(dev_inout, dev_inout, count * ncomp, nt, nop, (nccl_comm)ctx->comm,
                                   ctx->stream));
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
    AQC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host_inout, ctx->red_host, count * eb);
    return AQC_OK;
}

// Used by aqc_linklist_build: every rank hashes on ONE global grid, so that cell
// indices are comparable with the single-device run and halo particles never fall
// below the local r_min (the reference converts a negative float to unsigned there,
// LinkList.cl.in:74-77).  keys = 4 ordered-uint minima followed by 4 maxima.
int aqc_comm_minmax(aqc_ctx* ctx, uint32_t* keys)
{
    if (ctx->nranks <= 1 || !ctx->comm)
        return AQC_OK;
    AQC_NCCL(ctx, g_nccl.GroupStart());
    AQC_NCCL(ctx, g_nccl.AllReduce(keys, keys, 4, NCCL_UINT32, NCCL_MIN, (nccl_comm)ctx->comm, ctx->stream));
    AQC_NCCL(ctx, g_nccl.AllReduce(keys + 4, keys + 4, 4, NCCL_UINT32, NCCL_MAX, (nccl_comm)ctx->comm,
                                   ctx->stream));
    AQC_NCCL(ctx, g_nccl.GroupEnd());
    return AQC_OK;
}
