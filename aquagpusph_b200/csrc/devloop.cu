// devloop.cu -- device-side loops: a `while` of the tool pipeline as ONE CUDA graph whose WHILE
// conditional node holds a recorded pass over the loop body.
//
// What it replaces: the reference's Conditional / SetScalar / Reduction tools evaluate on the host
// (Conditional.cpp:85-96, SetScalar.cpp:146-195, Reduction.cpp:205-258): each sub-iteration of
// the midpoint scheme drains the queue at every reduction before the host can evaluate the
// relaxation / stop expressions and decide whether to enqueue another pass.  Here the scalars of
// the loop live in a table in device memory, the scalar tools are stack programs
// (include/aquasvm.h) run by a one-thread kernel between the body's kernels, and the program
// holding the loop condition calls cudaGraphSetConditional: the host launches one graph and
// synchronises once per loop, whatever the iteration count.
//
// Layout of the loop's device buffer:  [aqs_header 16 B][table][history: hist_rows x (16 + table)]
// Graph:  entry program (1 kernel node) -> WHILE node { recorded body }.
// The body is recorded with stream capture INTO the node's body graph
// (cudaStreamBeginCaptureToGraph, relaxed mode: allocations made by a launcher on first use are
// legal, anything that synchronises fails the capture and thereby the recording).
#include <chrono>
#include <stdlib.h>

#include "aqc_common.cuh"
#include "aquasvm.h"

struct aqc_loop {
    int table_bytes = 0, hist_rows = 0, max_ops = 0;
    char* dev = nullptr;          // header + table + history
    size_t dev_bytes = 0;
    char* host = nullptr;         // pinned mirror of the same (+ 16 bytes: max_iters staging)
    aqs_op* arena_dev = nullptr;  // programs
    aqs_op* arena_host = nullptr; // pinned staging
    int arena_used = 0;
    uint32_t* max_iters_dev = nullptr; // read by the programs' kernel (a launch-time value)
    cudaGraph_t graph = nullptr;
    cudaGraph_t body = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphConditionalHandle handle = 0;
    cudaEvent_t join_ev = nullptr; // aqc_loop_abort: joins a branch lane left inside the capture
    int arena_graph_from = 0;     // first op of the recorded programs (direct ones sit in front)
    bool recording = false, ready = false, body_has_cond = false;
    bool started = false;         // aqc_loop_start uploaded the table: aqc_loop_run(table_in = NULL) may follow
    uint64_t launches_at_begin = 0, body_launches = 0;
    int body_nodes = 0;
    double record_ms = 0.0, instantiate_ms = 0.0;
    std::chrono::steady_clock::time_point t_begin;
};

namespace {

__global__ void __launch_bounds__(32)
svm_kernel(const aqs_op* __restrict__ prog, int n, char* __restrict__ base, int table_bytes,
           int hist_rows, cudaGraphConditionalHandle handle, int has_cond,
           const uint32_t* __restrict__ max_iters)
{
    if (threadIdx.x != 0)
        return;
    aqs_header* hdr = (aqs_header*)base;
    char* tab = base + sizeof(aqs_header);
    char* hist = hist_rows > 0 ? tab + table_bytes : nullptr;
    const int cond = aqs_run(prog, n, tab, table_bytes, hdr, hist, hist_rows, *max_iters);
    if (has_cond)
        cudaGraphSetConditional(handle, (cond > 0 && !hdr->error) ? 1u : 0u);
}

bool has_setcond(const aqs_op* prog, int n)
{
    for (int k = 0; k < n; k++)
        if (prog[k].code == AQS_SETCOND || prog[k].code == AQS_RECOND)
            return true;
    return false;
}

void drop_graph(aqc_loop* L)
{
    if (L->exec)
        cudaGraphExecDestroy(L->exec);
    if (L->graph)
        cudaGraphDestroy(L->graph); // owns the body graph
    L->exec = nullptr;
    L->graph = nullptr;
    L->body = nullptr;
    L->ready = false;
}

// a recording that did not end well: nothing of it ran, although the library's bookkeeping
// (pair caches, watches, sync plans, the "clean" flags of the link-list scratch) saw the calls
void forget_recorded_state(aqc_ctx* ctx)
{
    aqc_pc_invalidate(ctx);
    for (aqc_watch& w : ctx->watches)
        w.dirty = true;
    for (aqc_sync_plan& pl : ctx->plans)
        pl.valid = false;
    ctx->sort_ghist_clean = false;
    ctx->minmax_clean = false;
}

} // namespace

extern "C" int aqc_loop_create(aqc_ctx* ctx, int table_bytes, int hist_rows, int max_ops, aqc_loop** out)
{
    if (!ctx || !out || table_bytes <= 0 || (table_bytes & 15) || hist_rows < 0 || max_ops <= 0)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_create: bad argument");
    *out = nullptr;
    aqc_loop* L = new aqc_loop();
    L->table_bytes = table_bytes;
    L->hist_rows = hist_rows;
    L->max_ops = max_ops;
    L->dev_bytes = sizeof(aqs_header) + (size_t)table_bytes + (size_t)hist_rows * (16 + (size_t)table_bytes);
    if (cudaMalloc(&L->dev, L->dev_bytes) != cudaSuccess ||
        cudaMallocHost(&L->host, L->dev_bytes + 16) != cudaSuccess ||
        cudaMalloc(&L->arena_dev, (size_t)max_ops * sizeof(aqs_op)) != cudaSuccess ||
        cudaMallocHost(&L->arena_host, (size_t)max_ops * sizeof(aqs_op)) != cudaSuccess ||
        cudaMalloc(&L->max_iters_dev, 16) != cudaSuccess) {
        cudaGetLastError();
        aqc_loop_destroy(ctx, L);
        return aqc_fail(ctx, AQC_ERR_CUDA, "aqc_loop_create: cannot allocate %zu bytes", L->dev_bytes);
    }
    *out = L;
    return AQC_OK;
}

extern "C" int aqc_loop_destroy(aqc_ctx* ctx, aqc_loop* L)
{
    if (!L)
        return AQC_OK;
    if (L->recording)
        aqc_loop_abort(ctx, L);
    if (ctx && ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    drop_graph(L);
    cudaFree(L->dev);
    cudaFreeHost(L->host);
    cudaFree(L->arena_dev);
    cudaFreeHost(L->arena_host);
    cudaFree(L->max_iters_dev);
    if (L->join_ev)
        cudaEventDestroy(L->join_ev);
    delete L;
    return AQC_OK;
}

extern "C" void* aqc_loop_table(aqc_loop* L) { return L ? L->dev + sizeof(aqs_header) : nullptr; }

extern "C" int aqc_loop_begin(aqc_ctx* ctx, aqc_loop* L, const aqs_op* entry, int n_entry)
{
    if (!ctx || !L || !entry || n_entry <= 0)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_begin: bad argument");
    if (ctx->recording)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_begin: a loop is already being recorded");
    if (ctx->lane != 0)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_begin: the branch lane is selected");
    if (!has_setcond(entry, n_entry))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_begin: the entry program holds no AQS_SETCOND");
    if (!L->started)
        L->arena_used = 0; // (programs run directly since aqc_loop_start stay where they are)
    if (L->arena_used + n_entry > L->max_ops)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_begin: program arena too small");
    L->t_begin = std::chrono::steady_clock::now();
    // the previous graph may still be referenced by nothing: aqc_loop_run synchronises
    drop_graph(L);
    const auto t_drop = std::chrono::steady_clock::now();
    L->body_has_cond = false;
    const int entry_at = L->arena_used;
    memcpy(L->arena_host + entry_at, entry, (size_t)n_entry * sizeof(aqs_op));
    L->arena_used += n_entry;
    L->arena_graph_from = entry_at;

    AQC_CUDA(ctx, cudaGraphCreate(&L->graph, 0));
    // the default value only matters if the entry program did not run: never
    AQC_CUDA(ctx, cudaGraphConditionalHandleCreate(&L->handle, L->graph, 0, cudaGraphCondAssignDefault));
    // entry program node
    const aqs_op* prog = L->arena_dev + entry_at;
    int n = n_entry, tb = L->table_bytes, hr = L->hist_rows, hc = 1;
    char* base = L->dev;
    const uint32_t* mi = L->max_iters_dev;
    void* kargs[] = { &prog, &n, &base, &tb, &hr, &L->handle, &hc, &mi };
    cudaKernelNodeParams kp = {};
    kp.func = (void*)svm_kernel;
    kp.gridDim = dim3(1);
    kp.blockDim = dim3(32);
    kp.sharedMemBytes = 0;
    kp.kernelParams = kargs;
    kp.extra = nullptr;
    cudaGraphNode_t entry_node = nullptr;
    AQC_CUDA(ctx, cudaGraphAddKernelNode(&entry_node, L->graph, nullptr, 0, &kp));
    cudaGraphNodeParams cp = {};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = L->handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t while_node = nullptr;
    AQC_CUDA(ctx, cudaGraphAddNode(&while_node, L->graph, &entry_node, 1, &cp));
    L->body = cp.conditional.phGraph_out[0];
    AQC_CUDA(ctx, cudaStreamBeginCaptureToGraph(ctx->stream, L->body, nullptr, nullptr, 0,
                                                cudaStreamCaptureModeRelaxed));
    L->recording = true;
    ctx->recording = L;
    L->launches_at_begin = ctx->launches;
    if (getenv("AQC_LOOP_DEBUG")) {
        const auto t_cap = std::chrono::steady_clock::now();
        fprintf(stderr, "aqc_loop_begin: drop %.3f ms, graph + capture start %.3f ms\n",
                std::chrono::duration<double, std::milli>(t_drop - L->t_begin).count(),
                std::chrono::duration<double, std::milli>(t_cap - t_drop).count());
    }
    return AQC_OK;
}

extern "C" int aqc_loop_svm(aqc_ctx* ctx, aqc_loop* L, const aqs_op* prog, int n)
{
    if (!ctx || !L || !prog || n <= 0)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_svm: bad argument");
    if (ctx->recording && ctx->recording != L)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_svm: another loop is being recorded");
    if (L->arena_used + n > L->max_ops)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_svm: program arena full (%d ops)", L->max_ops);
    memcpy(L->arena_host + L->arena_used, prog, (size_t)n * sizeof(aqs_op));
    if (!L->recording) {
        // direct: the program runs now (no graph, so no condition to set)
        AQC_CUDA(ctx, cudaMemcpyAsync(L->arena_dev + L->arena_used, L->arena_host + L->arena_used,
                                      (size_t)n * sizeof(aqs_op), cudaMemcpyHostToDevice, ctx->stream));
        svm_kernel<<<1, 32, 0, ctx->stream>>>(L->arena_dev + L->arena_used, n, L->dev, L->table_bytes,
                                              L->hist_rows, 0, 0, L->max_iters_dev);
        L->arena_used += n;
        AQC_LAUNCH_CHECK(ctx);
        return AQC_OK;
    }
    const int hc = has_setcond(prog, n) ? 1 : 0;
    svm_kernel<<<1, 32, 0, ctx->stream>>>(L->arena_dev + L->arena_used, n, L->dev, L->table_bytes,
                                          L->hist_rows, L->handle, hc, L->max_iters_dev);
    L->arena_used += n;
    L->body_has_cond |= hc != 0;
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

extern "C" int aqc_loop_abort(aqc_ctx* ctx, aqc_loop* L)
{
    if (!ctx || !L)
        return AQC_ERR_ARG;
    if (L->recording) {
        if (ctx->lane != 0)
            aqc_lane_select(ctx, 0);
        // a branch lane that joined the capture and was not waited for yet would leave the capture
        // "unjoined": join it, so that both streams leave capture mode cleanly
        if (ctx->lane1_made && ctx->parked.stream) {
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(ctx->parked.stream, &st) == cudaSuccess &&
                st == cudaStreamCaptureStatusActive) {
                if (!L->join_ev)
                    cudaEventCreateWithFlags(&L->join_ev, cudaEventDisableTiming);
                if (L->join_ev && cudaEventRecord(L->join_ev, ctx->parked.stream) == cudaSuccess)
                    cudaStreamWaitEvent(ctx->stream, L->join_ev, 0);
            }
            cudaGetLastError();
        }
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(ctx->stream, &g); // an invalidated capture reports its error here
        cudaGetLastError();
        L->recording = false;
        ctx->recording = nullptr;
        ctx->launches = L->launches_at_begin;
        forget_recorded_state(ctx);
    }
    drop_graph(L);
    return AQC_OK;
}

extern "C" int aqc_loop_end(aqc_ctx* ctx, aqc_loop* L)
{
    if (!ctx || !L)
        return AQC_ERR_ARG;
    if (!L->recording)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_end: the loop is not being recorded");
    if (ctx->lane != 0)
        aqc_lane_select(ctx, 0);
    if (!L->body_has_cond) {
        aqc_loop_abort(ctx, L);
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_end: the body never sets the loop condition");
    }
    cudaGraph_t g = nullptr;
    const auto t_e0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (getenv("AQC_LOOP_DEBUG"))
        fprintf(stderr, "aqc_loop_end: body recorded in %.3f ms since begin, end capture %.3f ms\n",
                std::chrono::duration<double, std::milli>(t_e0 - L->t_begin).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_e0).count());
    L->recording = false;
    ctx->recording = nullptr;
    L->body_launches = ctx->launches - L->launches_at_begin;
    ctx->launches = L->launches_at_begin; // nothing ran yet: aqc_loop_run counts what does
    if (e != cudaSuccess) {
        cudaGetLastError();
        forget_recorded_state(ctx);
        drop_graph(L);
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_end: the body cannot be recorded (%s)",
                        cudaGetErrorString(e));
    }
    const auto t1 = std::chrono::steady_clock::now();
    L->record_ms = std::chrono::duration<double, std::milli>(t1 - L->t_begin).count();
    size_t nn = 0;
    cudaGraphGetNodes(L->body, nullptr, &nn);
    L->body_nodes = (int)nn;
    e = cudaGraphInstantiate(&L->exec, L->graph, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        forget_recorded_state(ctx);
        drop_graph(L);
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_end: cudaGraphInstantiate failed (%s)",
                        cudaGetErrorString(e));
    }
    L->instantiate_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    // the programs travel behind whatever is queued, in front of the launch
    AQC_CUDA(ctx, cudaMemcpyAsync(L->arena_dev + L->arena_graph_from, L->arena_host + L->arena_graph_from,
                                  (size_t)(L->arena_used - L->arena_graph_from) * sizeof(aqs_op),
                                  cudaMemcpyHostToDevice, ctx->stream));
    L->ready = true;
    return AQC_OK;
}

extern "C" int aqc_loop_start(aqc_ctx* ctx, aqc_loop* L, const void* table_in, uint32_t max_iters)
{
    if (!ctx || !L || !table_in)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_start: bad argument");
    if (ctx->recording)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_start: a loop is being recorded");
    const size_t head = sizeof(aqs_header) + (size_t)L->table_bytes;
    // (the pinned mirror is free: the previous run synchronised)
    memset(L->host, 0, sizeof(aqs_header));
    memcpy(L->host + sizeof(aqs_header), table_in, (size_t)L->table_bytes);
    uint32_t* mi_host = (uint32_t*)(L->host + L->dev_bytes); // 16 spare bytes behind the mirror
    *mi_host = max_iters;
    AQC_CUDA(ctx, cudaMemcpyAsync(L->max_iters_dev, mi_host, 4, cudaMemcpyHostToDevice, ctx->stream));
    AQC_CUDA(ctx, cudaMemcpyAsync(L->dev, L->host, head, cudaMemcpyHostToDevice, ctx->stream));
    L->arena_used = 0;
    L->started = true;
    L->ready = false; // a body recorded for other values is not the one to launch
    return AQC_OK;
}

extern "C" int aqc_loop_run(aqc_ctx* ctx, aqc_loop* L, const void* table_in, uint32_t max_iters,
                            aqs_header* hdr_out, void* table_out, void* hist_out)
{
    if (!ctx || !L)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_loop_run: bad argument");
    if (ctx->recording)
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_run: a loop is being recorded");
    if (table_in) {
        if (!L->ready)
            return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_run: no recorded body");
        const bool ready = L->ready;
        const int used = L->arena_used;
        const int rc = aqc_loop_start(ctx, L, table_in, max_iters);
        L->ready = ready;      // the same recording, other initial values
        L->arena_used = used;
        if (rc)
            return rc;
    } else if (!L->started) {
        return aqc_fail(ctx, AQC_ERR_STATE, "aqc_loop_run: neither a table nor aqc_loop_start");
    }
    const size_t head = sizeof(aqs_header) + (size_t)L->table_bytes;
    if (L->ready)
        AQC_CUDA(ctx, cudaGraphLaunch(L->exec, ctx->stream));
    AQC_CUDA(ctx, cudaMemcpyAsync(L->host, L->dev, L->dev_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    AQC_SYNC(ctx);
    L->started = false;
    const aqs_header* h = (const aqs_header*)L->host;
    const bool ran = L->ready;
    // a graph recorded for the table aqc_loop_start uploaded is used once (the scalars baked into
    // its kernels belong to this time step): released now, while the device is idle anyway --
    // destroying an executable graph can synchronise the device, which in front of the next
    // recording would wait for the pass that is running
    if (!table_in)
        drop_graph(L);
    if (ran) // the entry program + the passes the graph made: one per condition that came out true
        ctx->launches += 1 + (uint64_t)h->iters * L->body_launches; // (a first pass run directly counted itself)
    if (hdr_out)
        *hdr_out = *h;
    if (table_out)
        memcpy(table_out, L->host + sizeof(aqs_header), (size_t)L->table_bytes);
    if (hist_out && L->hist_rows)
        memcpy(hist_out, L->host + head, (size_t)L->hist_rows * (16 + (size_t)L->table_bytes));
    return AQC_OK;
}

extern "C" int aqc_loop_stats(const aqc_loop* L, int* body_nodes, double* record_ms, double* instantiate_ms)
{
    if (!L)
        return AQC_ERR_ARG;
    if (body_nodes)
        *body_nodes = L->body_nodes;
    if (record_ms)
        *record_ms = L->record_ms;
    if (instantiate_ms)
        *instantiate_ms = L->instantiate_ms;
    return AQC_OK;
}
