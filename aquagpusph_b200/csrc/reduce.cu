// reduce.cu -- Reduction tool (aquagpusph/CalcServer/Reduction.cpp:143-258,
// Reduction.cl.in:35-66) for the operations the presets use: sum, min, max over
// float / uint / int / vec2 / vec4 arrays.
//
// The reference runs a log_wg(N) cascade of tree kernels and reads the result
// back through an event callback.  Here: one grid-stride pass (fixed grid =>
// fixed summation order => run-to-run deterministic), warp-shuffle and
// shared-memory block reduction, per-CTA partials, and a second single-CTA pass.
// min/max are order independent, hence bit-exact against the reference; sums
// differ from the reference's tree order by fp32 rounding only (the reference's
// own result is device dependent, Reduction.cpp:337-373).
#include <math.h>

#include "aqc_common.cuh"

namespace {

template <int OP> struct Op;
template <> struct Op<AQC_OP_SUM> {
    template <typename T> __device__ static T apply(T a, T b) { return a + b; }
    __device__ static float idf() { return 0.f; }
    __device__ static uint32_t idu() { return 0u; }
    __device__ static int idi() { return 0; }
};
template <> struct Op<AQC_OP_MIN> {
    __device__ static float apply(float a, float b) { return fminf(a, b); }
    __device__ static uint32_t apply(uint32_t a, uint32_t b) { return a < b ? a : b; }
    __device__ static int apply(int a, int b) { return a < b ? a : b; }
    __device__ static float idf() { return INFINITY; }
    __device__ static uint32_t idu() { return 0xFFFFFFFFu; }
    __device__ static int idi() { return 0x7FFFFFFF; }
};
template <> struct Op<AQC_OP_MAX> {
    __device__ static float apply(float a, float b) { return fmaxf(a, b); }
    __device__ static uint32_t apply(uint32_t a, uint32_t b) { return a > b ? a : b; }
    __device__ static int apply(int a, int b) { return a > b ? a : b; }
    __device__ static float idf() { return -INFINITY; }
    __device__ static uint32_t idu() { return 0u; }
    __device__ static int idi() { return (int)0x80000000; }
};

template <typename T, int OP> __device__ T ident();
template <> __device__ float ident<float, 0>() { return Op<0>::idf(); }
template <> __device__ float ident<float, 1>() { return Op<1>::idf(); }
template <> __device__ float ident<float, 2>() { return Op<2>::idf(); }
template <> __device__ uint32_t ident<uint32_t, 0>() { return Op<0>::idu(); }
template <> __device__ uint32_t ident<uint32_t, 1>() { return Op<1>::idu(); }
template <> __device__ uint32_t ident<uint32_t, 2>() { return Op<2>::idu(); }
template <> __device__ int ident<int, 0>() { return Op<0>::idi(); }
template <> __device__ int ident<int, 1>() { return Op<1>::idi(); }
template <> __device__ int ident<int, 2>() { return Op<2>::idi(); }

// NC interleaved components of scalar type T per element (1, 2 or 4)
template <typename T, int NC, int OP>
__global__ void __launch_bounds__(256)
reduce_kernel(const T* __restrict__ in, size_t n, T* __restrict__ out)
{
    T acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++)
        acc[c] = ident<T, OP>();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] = Op<OP>::apply(acc[c], in[i * NC + c]);
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc[c] = Op<OP>::apply(acc[c], __shfl_down_sync(0xffffffffu, acc[c], o));
    __shared__ T sm[8][NC];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
#pragma unroll
        for (int c = 0; c < NC; c++)
            sm[w][c] = acc[c];
    __syncthreads();
    if (threadIdx.x < NC) {
        T a = sm[0][threadIdx.x];
        for (int k = 1; k < 8; k++)
            a = Op<OP>::apply(a, sm[k][threadIdx.x]);
        out[(size_t)blockIdx.x * NC + threadIdx.x] = a;
    }
}

template <typename T, int NC, int OP>
int run(aqc_ctx* ctx, const void* in, size_t n, void* out_dev, void* out_host)
{
    unsigned grid = aqc_blocks(n ? n : 1, 256 * 4);
    const unsigned cap = (unsigned)ctx->sm_count * 4;
    if (grid > cap)
        grid = cap;
    const size_t need = ((size_t)grid + 1) * NC * sizeof(T);
    if (need > ctx->red_cap) {
        if (ctx->red_dev)
            AQC_CUDA(ctx, cudaFree(ctx->red_dev));
        ctx->red_dev = nullptr;
        ctx->red_cap = 0;
        AQC_CUDA(ctx, cudaMalloc(&ctx->red_dev, need + 4096));
        ctx->red_cap = need + 4096;
    }
    T* partial = (T*)ctx->red_dev;
    T* final_dev = out_dev ? (T*)out_dev : partial + (size_t)grid * NC;
    reduce_kernel<T, NC, OP><<<grid, 256, 0, ctx->stream>>>((const T*)in, n, partial);
    AQC_LAUNCH_CHECK(ctx);
    reduce_kernel<T, NC, OP><<<1, 256, 0, ctx->stream>>>(partial, grid, final_dev);
    AQC_LAUNCH_CHECK(ctx);
    if (out_host) {
        AQC_CUDA(ctx, cudaMemcpyAsync(ctx->red_host, final_dev, NC * sizeof(T),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        AQC_SYNC(ctx);
        memcpy(out_host, ctx->red_host, NC * sizeof(T));
    }
    return AQC_OK;
}

template <typename T, int NC>
int run_op(aqc_ctx* ctx, int op, const void* in, size_t n, void* od, void* oh)
{
    switch (op) {
        case AQC_OP_SUM: return run<T, NC, AQC_OP_SUM>(ctx, in, n, od, oh);
        case AQC_OP_MIN: return run<T, NC, AQC_OP_MIN>(ctx, in, n, od, oh);
        case AQC_OP_MAX: return run<T, NC, AQC_OP_MAX>(ctx, in, n, od, oh);
    }
    return aqc_fail(ctx, AQC_ERR_ARG, "aqc_reduce: unknown op %d", op);
}

} // namespace

extern "C" int aqc_reduce(aqc_ctx* ctx, int op, int type, const void* in, size_t n, void* out_dev,
                          void* out_host)
{
    if (!ctx || (!in && n))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_reduce: NULL input");
    aqc_pc_touch(ctx, out_dev, type == AQC_T_VEC4 ? 16 : type == AQC_T_VEC2 ? 8 : 4);
    switch (type) {
        case AQC_T_F32: return run_op<float, 1>(ctx, op, in, n, out_dev, out_host);
        case AQC_T_U32: return run_op<uint32_t, 1>(ctx, op, in, n, out_dev, out_host);
        case AQC_T_I32: return run_op<int, 1>(ctx, op, in, n, out_dev, out_host);
        case AQC_T_VEC2: return run_op<float, 2>(ctx, op, in, n, out_dev, out_host);
        case AQC_T_VEC4: return run_op<float, 4>(ctx, op, in, n, out_dev, out_host);
    }
    return aqc_fail(ctx, AQC_ERR_ARG, "aqc_reduce: unknown type %d", type);
}
