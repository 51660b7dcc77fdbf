// aqc_api.cu -- context, memory, fill, events and the kernel registry of the
// C-ABI declared in include/aquacuda.h.
#include <stdarg.h>
#include <stdlib.h>

#include "aqc_common.cuh"

int aqc_fail(aqc_ctx* ctx, int code, const char* fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
        // a rank that hit a device fault cannot keep step with its peers: give the communicator
        // up now, so that this process can leave and the peers' bounded waits end (mpi.cu)
        if (code == AQC_ERR_CUDA && ctx->comm && !ctx->recording) // (a failed recording ran nothing)
            aqc_comm_abort(ctx);
    }
    return code;
}

static char g_create_err[512] = "";

extern "C" int aqc_ctx_create(int device, aqc_ctx** out)
{
    if (!out)
        return AQC_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "no usable CUDA device %d (count=%d, %s); libaquacuda has no "
                 "CPU fallback", device, n,
                 e == cudaSuccess ? "ok" : cudaGetErrorString(e));
        return AQC_ERR_CUDA;
    }
    aqc_ctx* ctx = new aqc_ctx();
    ctx->device = device;
    // (the greatest priority: the branch lane of aqc_lane_select runs one class below)
    int prio_least = 0, prio_greatest = 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cannot initialise device %d", device);
        delete ctx;
        return AQC_ERR_CUDA;
    }
    ctx->own_stream = true;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaMalloc(&ctx->minmax_dev, 8 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMallocHost(&ctx->minmax_host, 8 * sizeof(float)) != cudaSuccess ||
        cudaMallocHost(&ctx->red_host, 64) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cannot allocate context scratch");
        delete ctx;
        return AQC_ERR_CUDA;
    }
    *out = ctx;
    return AQC_OK;
}

extern "C" void aqc_ctx_destroy(aqc_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    aqc_comm_destroy(ctx); // (drains the stream with a deadline while a communicator is live)
    if (!ctx->comm_dead)   // (after an abort a kernel of this rank may never end)
        cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 2; k++) {
        cudaFree(ctx->sort_keys[k]);
        cudaFree(ctx->sort_vals[k]);
    }
    cudaFree(ctx->sort_hist);
    cudaFree(ctx->minmax_dev);
    cudaFree(ctx->cell_cls);
    cudaFree(ctx->pack_rows);
    if (ctx->lane1_made) {
        if (ctx->lane != 0)
            aqc_lane_select(ctx, 0);
        cudaStreamSynchronize(ctx->parked.stream);
        cudaStreamDestroy(ctx->parked.stream);
        cudaFree(ctx->parked.cell_cls);
        cudaFree(ctx->parked.pack_rows);
    }
    for (aqc_pair_cache* c : { &ctx->pc, &ctx->pcr }) {
        cudaFree(c->masks);
        cudaFree(c->chunks);
        cudaFree(c->cnt);
        cudaFree(c->pass_tab);
        cudaFree(c->ctl);
        cudaFreeHost(c->ctl_host);
    }
    cudaFreeHost(ctx->minmax_host);
    cudaFree(ctx->red_dev);
    cudaFreeHost(ctx->red_host);
    if (ctx->side) {
        cudaStreamSynchronize(ctx->side);
        cudaStreamDestroy(ctx->side);
        cudaEventDestroy(ctx->side_fork_ev);
    }
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* aqc_last_error(const aqc_ctx* ctx)
{
    return ctx ? ctx->err : g_create_err;
}

extern "C" int aqc_set_stream(aqc_ctx* ctx, void* s)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (ctx->lane != 0)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_set_stream: the branch lane is selected");
    AQC_SYNC(ctx);
    if (ctx->own_stream) {
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    if (s) {
        ctx->stream = (cudaStream_t)s;
    } else {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        AQC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, greatest));
        ctx->own_stream = true;
    }
    return AQC_OK;
}

extern "C" void* aqc_get_stream(aqc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int aqc_sync(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    AQC_SYNC(ctx);
    return AQC_OK;
}

// ---- lanes: a second in-order queue for tools that do not depend on what the first is doing ----
extern "C" int aqc_lane_select(aqc_ctx* ctx, int lane)
{
    if (!ctx || (lane != 0 && lane != 1))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_lane_select: bad argument");
    if (lane == ctx->lane)
        return AQC_OK;
    if (!ctx->lane1_made) {
        // one step below the context's stream where the device has priorities: the branch usually
        // carries the long sweep, and the short kernels of the main lane should find their way
        // in between its CTAs (AQC_LANE1_PRIORITY overrides; lower number = served first)
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        int prio = least;
        if (const char* e = getenv("AQC_LANE1_PRIORITY"))
            prio = atoi(e);
        AQC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->parked.stream, cudaStreamNonBlocking, prio));
        ctx->lane1_made = true;
    }
    aqc_ctx::lane_state cur;
    cur.stream = ctx->stream;
    cur.cell_cls = ctx->cell_cls;
    cur.cell_cls_cap = ctx->cell_cls_cap;
    cur.pack_rows = ctx->pack_rows;
    cur.pack_cap = ctx->pack_cap;
    ctx->stream = ctx->parked.stream;
    ctx->cell_cls = ctx->parked.cell_cls;
    ctx->cell_cls_cap = ctx->parked.cell_cls_cap;
    ctx->pack_rows = ctx->parked.pack_rows;
    ctx->pack_cap = ctx->parked.pack_cap;
    ctx->parked = cur;
    ctx->lane = lane;
    return AQC_OK;
}

extern "C" int aqc_lane_event(aqc_ctx* ctx, void** ev)
{
    if (!ctx || !ev)
        return AQC_ERR_ARG;
    if (!*ev) {
        cudaEvent_t e;
        AQC_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        *ev = (void*)e;
    }
    AQC_CUDA(ctx, cudaEventRecord((cudaEvent_t)*ev, ctx->stream));
    return AQC_OK;
}

extern "C" int aqc_lane_wait(aqc_ctx* ctx, void* ev)
{
    if (!ctx || !ev)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, (cudaEvent_t)ev, 0));
    return AQC_OK;
}

// ---- side stream of the savers (Particles.cpp:243-323) ---------------------------------------
extern "C" int aqc_side_fork(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (!ctx->side) {
        AQC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
        AQC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->side_fork_ev, cudaEventDisableTiming));
    }
    AQC_CUDA(ctx, cudaEventRecord(ctx->side_fork_ev, ctx->stream));
    AQC_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->side_fork_ev, 0));
    return AQC_OK;
}

extern "C" int aqc_memcpy_d2h_side(aqc_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx || !ctx->side)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_memcpy_d2h_side: aqc_side_fork first");
    if (!bytes)
        return AQC_OK;
    if (!dst || !src)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_memcpy_d2h_side: NULL pointer");
    AQC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->side));
    return AQC_OK;
}

extern "C" int aqc_side_record(aqc_ctx* ctx, void* ev)
{
    if (!ctx || !ctx->side || !ev)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_side_record: aqc_side_fork first, ev != NULL");
    AQC_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->side));
    return AQC_OK;
}

extern "C" int aqc_side_wait(aqc_ctx* ctx, void* ev)
{
    if (!ctx || !ev)
        return AQC_ERR_ARG;
    // (no aqc_fail here: ctx->err belongs to the thread that drives the context)
    if (cudaSetDevice(ctx->device) != cudaSuccess)
        return AQC_ERR_CUDA;
    return cudaEventSynchronize((cudaEvent_t)ev) == cudaSuccess ? AQC_OK : AQC_ERR_CUDA;
}

extern "C" uint64_t aqc_launch_count(const aqc_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int aqc_device_sm_count(const aqc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// CalcServer.cpp:245-257: snprintf("%#G") + "f", parsed back by the compiler.
extern "C" float aqc_define_round6(float value)
{
    char s[128];
    snprintf(s, sizeof(s), "%#G", (double)value);
    return strtof(s, nullptr);
}

extern "C" int aqc_set_defs(aqc_ctx* ctx, const aqc_defs* defs)
{
    if (!ctx || !defs || (defs->dims != 2 && defs->dims != 3))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_set_defs: dims must be 2 or 3");
    ctx->defs = *defs;
    return AQC_OK;
}

static float parse_define_float(const char* v)
{
    std::string s(v ? v : "");
    while (!s.empty() && (s.back() == 'f' || s.back() == 'F' || isspace((unsigned char)s.back())))
        s.pop_back();
    return strtof(s.c_str(), nullptr);
}

extern "C" int aqc_set_define(aqc_ctx* ctx, const char* name, const char* value)
{
    if (!ctx || !name)
        return AQC_ERR_ARG;
    const std::string n(name), v(value ? value : "");
    if (n == "H") ctx->defs.H = parse_define_float(value);
    else if (n == "CONW") ctx->defs.CONW = parse_define_float(value);
    else if (n == "CONF") ctx->defs.CONF = parse_define_float(value);
    else if (n == "SUPPORT") ctx->defs.SUPPORT = parse_define_float(value);
    else if (n == "DIMS") ctx->defs.DIMS = parse_define_float(value);
    else if (n == "__DR_FACTOR__") { ctx->dr_factor = parse_define_float(value); ctx->has_dr_factor = true; }
    else if (n == "__MIN_BOUND_DIST__") { ctx->min_bound_dist = parse_define_float(value); ctx->has_min_bound_dist = true; }
    else if (n == "__ELASTIC_FACTOR__") ctx->elastic_factor = parse_define_float(value);
    else if (n == "TSCHEME_ADAMS_BASHFORTH_STEPS") ctx->ab_steps = (unsigned)strtoul(v.c_str(), nullptr, 10); // "5u"
    else if (n == "KERNEL_NAME") {
        if (v != "Wendland")
            return aqc_fail(ctx, AQC_ERR_ARG, "KERNEL_NAME=%s: only the Wendland kernel is built",
                            v.c_str());
    } else if (n == "__LAP_FORMULATION__") {
        if (v == "2" || v == "__LAP_MORRIS__")
            ctx->lap_morris = true; // cfd/Interactions.cl runs PInteractionsMorris, nothing is fused
        else if (v == "1" || v == "__LAP_MONAGHAN__")
            ctx->lap_morris = false;
        else
            return aqc_fail(ctx, AQC_ERR_ARG,
                            "__LAP_FORMULATION__=%s: __LAP_MONAGHAN__ and __LAP_MORRIS__ are built", v.c_str());
    } else
        return 1;
    return AQC_OK;
}

extern "C" int aqc_alloc(aqc_ctx* ctx, size_t bytes, void** dptr)
{
    if (!ctx || !dptr)
        return AQC_ERR_ARG;
    *dptr = nullptr;
    if (!bytes)
        bytes = 16;
    AQC_CUDA(ctx, cudaMalloc(dptr, bytes));
    return AQC_OK;
}

extern "C" int aqc_free(aqc_ctx* ctx, void* dptr)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (dptr) {
        aqc_pc_touch(ctx, dptr, 1);
        AQC_SYNC(ctx);
        AQC_CUDA(ctx, cudaFree(dptr));
    }
    return AQC_OK;
}

extern "C" int aqc_host_alloc(aqc_ctx* ctx, size_t bytes, void** hptr)
{
    if (!ctx || !hptr)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 16));
    return AQC_OK;
}

extern "C" int aqc_host_free(aqc_ctx* ctx, void* hptr)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (hptr)
        AQC_CUDA(ctx, cudaFreeHost(hptr));
    return AQC_OK;
}

extern "C" int aqc_memcpy_h2d(aqc_ctx* ctx, void* dst, const void* src, size_t bytes, int blocking)
{
    if (!ctx)
        return AQC_ERR_ARG;
    aqc_pc_touch(ctx, dst, bytes);
    AQC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (blocking)
        AQC_SYNC(ctx);
    return AQC_OK;
}

extern "C" int aqc_memcpy_d2h(aqc_ctx* ctx, void* dst, const void* src, size_t bytes, int blocking)
{
    if (!ctx)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (blocking)
        AQC_SYNC(ctx);
    return AQC_OK;
}

extern "C" int aqc_memcpy_d2d(aqc_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx)
        return AQC_ERR_ARG;
    aqc_pc_touch(ctx, dst, bytes);
    AQC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return AQC_OK;
}

// ---- Set tool -------------------------------------------------------------
template <typename T>
__global__ void fill_kernel(T* __restrict__ p, size_t n, T v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void fill64_kernel(uint4* __restrict__ p, size_t n4, uint4 a, uint4 b, uint4 c, uint4 d)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        const unsigned k = (unsigned)(i & 3);
        p[i] = k == 0 ? a : (k == 1 ? b : (k == 2 ? c : d));
    }
}

template <typename T>
static int fill_launch(aqc_ctx* ctx, void* p, size_t n, const void* value)
{
    T v;
    memcpy(&v, value, sizeof(T));
    unsigned bs = 256;
    unsigned grid = aqc_blocks(n, bs);
    unsigned cap = (unsigned)ctx->sm_count * 16;
    if (grid > cap)
        grid = cap;
    fill_kernel<T><<<grid, bs, 0, ctx->stream>>>((T*)p, n, v);
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

extern "C" int aqc_fill(aqc_ctx* ctx, void* dptr, size_t n, size_t elem_bytes, const void* value)
{
    if (!ctx || !value)
        return AQC_ERR_ARG;
    if (!n)
        return AQC_OK;
    aqc_pc_touch(ctx, dptr, n * elem_bytes);
    switch (elem_bytes) {
        case 4: return fill_launch<uint32_t>(ctx, dptr, n, value);
        case 8: return fill_launch<uint2>(ctx, dptr, n, value);
        case 16: return fill_launch<uint4>(ctx, dptr, n, value);
        case 64: {
            // matrix (float16): four uint4 per element
            uint4 v[4];
            memcpy(v, value, 64);
            unsigned grid = aqc_blocks(n * 4, 256);
            unsigned cap = (unsigned)ctx->sm_count * 16;
            fill64_kernel<<<grid > cap ? cap : grid, 256, 0, ctx->stream>>>(
                (uint4*)dptr, n * 4, v[0], v[1], v[2], v[3]);
            AQC_LAUNCH_CHECK(ctx);
            return AQC_OK;
        }
        default:
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_fill: unsupported element size %zu", elem_bytes);
    }
}

// ---- events ---------------------------------------------------------------
extern "C" int aqc_event_create(aqc_ctx* ctx, void** ev)
{
    if (!ctx || !ev)
        return AQC_ERR_ARG;
    cudaEvent_t e;
    AQC_CUDA(ctx, cudaEventCreate(&e));
    *ev = (void*)e;
    return AQC_OK;
}
extern "C" int aqc_event_destroy(aqc_ctx* ctx, void* ev)
{
    if (!ctx)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaEventDestroy((cudaEvent_t)ev));
    return AQC_OK;
}
extern "C" int aqc_event_record(aqc_ctx* ctx, void* ev)
{
    if (!ctx)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->stream));
    return AQC_OK;
}
extern "C" int aqc_event_sync(aqc_ctx* ctx, void* ev)
{
    if (!ctx)
        return AQC_ERR_ARG;
    if (ctx->comm)
        return aqc_comm_wait(ctx, (cudaEvent_t)ev);
    AQC_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)ev));
    return AQC_OK;
}
extern "C" int aqc_event_elapsed_ms(aqc_ctx* ctx, void* a, void* b, float* ms)
{
    if (!ctx || !ms)
        return AQC_ERR_ARG;
    AQC_CUDA(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return AQC_OK;
}

// ---- kernel registry --------------------------------------------------------
std::vector<aqc_kernel_entry>& aqc_registry()
{
    static std::vector<aqc_kernel_entry> reg;
    return reg;
}

// Presets name scripts as "../Scripts/cfd/Interactions.cl" (cfd.xml:60) or with
// an absolute resources path; keep what follows the last "Scripts/".
static const char* strip_script_path(const char* p)
{
    const char* best = p;
    for (const char* s = p; (s = strstr(s, "Scripts/")) != nullptr; s += 8)
        best = s + 8;
    return best;
}

extern "C" int aqc_kernel_lookup(const char* script_path, const char* entry, int dims)
{
    if (!script_path)
        return AQC_ERR_ARG;
    const char* rel = strip_script_path(script_path);
    const char* ent = (entry && *entry) ? entry : "entry"; // State.cpp:1058-1059
    auto& reg = aqc_registry();
    // (run-time scripts are compiled per problem -- its definitions are baked in -- and are only
    // handed out by aqc_script_compile)
    for (size_t k = 0; k < reg.size(); k++)
        if (!reg[k].jit && !strcmp(reg[k].script, rel) && !strcmp(reg[k].entry, ent) &&
            (reg[k].dims == 0 || reg[k].dims == dims))
            return (int)k;
    // case-local scripts (outside resources/Scripts) are registered by file name
    const char* base = strrchr(rel, '/');
    base = base ? base + 1 : rel;
    for (size_t k = 0; k < reg.size(); k++)
        if (!reg[k].jit && !strchr(reg[k].script, '/') && !strcmp(reg[k].script, base) &&
            !strcmp(reg[k].entry, ent) && (reg[k].dims == 0 || reg[k].dims == dims))
            return (int)k;
    return AQC_ERR_NOKERNEL;
}

extern "C" int aqc_kernel_count(void) { return (int)aqc_registry().size(); }

extern "C" const char* aqc_kernel_name(int id)
{
    static thread_local std::string s;
    auto& reg = aqc_registry();
    if (id < 0 || id >= (int)reg.size())
        return nullptr;
    s = std::string(reg[id].script) + "::" + reg[id].entry;
    return s.c_str();
}

extern "C" int aqc_kernel_nargs(int id)
{
    auto& reg = aqc_registry();
    if (id < 0 || id >= (int)reg.size())
        return AQC_ERR_ARG;
    return (int)reg[id].args.size();
}

extern "C" const aqc_arg_info* aqc_kernel_args(int id)
{
    auto& reg = aqc_registry();
    if (id < 0 || id >= (int)reg.size())
        return nullptr;
    return reg[id].args.data();
}

extern "C" uint64_t aqc_kernel_dev_scalars(int id)
{
    auto& reg = aqc_registry();
    return (id < 0 || id >= (int)reg.size()) ? 0 : reg[id].dev_mask;
}

extern "C" int aqc_launch_ex(aqc_ctx* ctx, int id, size_t n, void* const* args, int nargs,
                             const void* const* dev_scalars)
{
    if (!ctx)
        return AQC_ERR_ARG;
    auto& reg = aqc_registry();
    if (id < 0 || id >= (int)reg.size())
        return aqc_fail(ctx, AQC_ERR_NOKERNEL, "aqc_launch_ex: bad kernel id %d", id);
    if (dev_scalars)
        for (int k = 0; k < nargs && k < (int)reg[id].args.size(); k++)
            if (dev_scalars[k] && (k >= 64 || !((reg[id].dev_mask >> k) & 1)))
                return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch_ex(%s::%s): argument %d (%s) cannot be "
                                "read from device memory", reg[id].script, reg[id].entry, k,
                                reg[id].args[k].name);
    ctx->dev_scalars = dev_scalars;
    const int rc = aqc_launch(ctx, id, n, args, nargs);
    ctx->dev_scalars = nullptr;
    return rc;
}

extern "C" int aqc_launch(aqc_ctx* ctx, int id, size_t n, void* const* args, int nargs)
{
    if (!ctx)
        return AQC_ERR_ARG;
    auto& reg = aqc_registry();
    if (id < 0 || id >= (int)reg.size())
        return aqc_fail(ctx, AQC_ERR_NOKERNEL, "aqc_launch: bad kernel id %d", id);
    if (nargs != (int)reg[id].args.size())
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch(%s::%s): expected %zu args, got %d",
                        reg[id].script, reg[id].entry, reg[id].args.size(), nargs);
    for (int k = 0; k < nargs; k++)
        if (!args[k])
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch(%s::%s): argument %d (%s) is NULL",
                            reg[id].script, reg[id].entry, k, reg[id].args[k].name);
    if (!n)
        return AQC_OK;
    for (int k = 0; k < nargs; k++) // what the kernel may write: one element per work-item
        if (reg[id].args[k].kind == AQC_ARG_ARRAY_OUT)
            aqc_pc_touch(ctx, args[k], n * aqc_type_bytes(reg[id].args[k].type, ctx->defs.dims));
    if (reg[id].jit)
        return aqc_script_launch(ctx, reg[id], n, args);
    return reg[id].fn(ctx, n, args);
}

// ---- write watches ---------------------------------------------------------------------------
extern "C" int aqc_watch_create(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    ctx->watches.emplace_back();
    return (int)ctx->watches.size() - 1;
}

extern "C" int aqc_watch_dirty(const aqc_ctx* ctx, int watch)
{
    if (!ctx || watch < 0 || watch >= (int)ctx->watches.size())
        return 1;
    return ctx->watches[watch].dirty ? 1 : 0;
}

extern "C" int aqc_watch_reset(aqc_ctx* ctx, int watch, int n, const void* const* ptrs, const size_t* bytes)
{
    if (!ctx || watch < 0 || watch >= (int)ctx->watches.size() || (n > 0 && (!ptrs || !bytes)))
        return AQC_ERR_ARG;
    aqc_watch& w = ctx->watches[watch];
    w.ranges.clear();
    for (int k = 0; k < n; k++)
        if (ptrs[k])
            w.ranges.emplace_back((const char*)ptrs[k], bytes[k]);
    w.dirty = w.ranges.empty();
    return AQC_OK;
}

// ---- pair-mask cache of the neighbour sweeps (sweep.cuh, S3Cache) ------------------------
extern "C" int aqc_pairs_cache_enable(aqc_ctx* ctx, int on)
{
    if (!ctx)
        return AQC_ERR_ARG;
    for (aqc_pair_cache* c : { &ctx->pc, &ctx->pcr }) {
        c->enabled = on != 0;
        c->valid = false;
        c->served = c->poor_streak = c->cooldown = 0; // (and the pay-off history)
    }
    return AQC_OK;
}

extern "C" int aqc_pairs_cache_invalidate(aqc_ctx* ctx)
{
    if (!ctx)
        return AQC_ERR_ARG;
    aqc_pc_invalidate(ctx);
    return AQC_OK;
}

static int pc_stats(const aqc_pair_cache& c, uint64_t* builds, uint64_t* hits, uint64_t* bytes)
{
    if (builds)
        *builds = c.builds;
    if (hits)
        *hits = c.hits;
    if (bytes)
        *bytes = c.lists ? (uint64_t)c.chunks_bytes + (uint64_t)c.cap_rounds * 7 * 32
                         : (uint64_t)c.cap_rounds * AQC_PC_ROUND_BYTES;
    return AQC_OK;
}
extern "C" int aqc_pairs_cache_stats(const aqc_ctx* ctx, uint64_t* builds, uint64_t* hits, uint64_t* bytes)
{
    return ctx ? pc_stats(ctx->pc, builds, hits, bytes) : AQC_ERR_ARG;
}
extern "C" int aqc_pairs_cache_stats_remote(const aqc_ctx* ctx, uint64_t* builds, uint64_t* hits, uint64_t* bytes)
{
    return ctx ? pc_stats(ctx->pcr, builds, hits, bytes) : AQC_ERR_ARG;
}
