// peak.cu -- measured FP32 (non-tensor) throughput of the device, the denominator of the
// "% of FP32 peak" figures of the neighbour sweeps.
//
// BASELINE.md section 2 asks for an FMA micro-benchmark before any such percentage is quoted:
// MEASURED_PEAKS.json (driver-written) holds the HBM copy bandwidth and the bf16 tensor rate only,
// and the sweeps use neither tensor cores nor much bandwidth.  Two kernels of independent
// register-to-register FMA chains, timed with CUDA events on the context's stream:
//   * scalar   fma.rn.f32      (SASS FFMA),  2 flop per lane and instruction
//   * packed   fma.rn.f32x2    (SASS FFMA2), 4 flop per lane and instruction -- sm_100's packed
//     fp32, which the candidate filter of the sweep engines uses
#include "aqc_common.cuh"

namespace {

constexpr int PEAK_ILP = 8;       // independent chains per thread
constexpr int PEAK_ITERS = 2048;  // FMAs per chain and launch (x 4 in the unrolled body)

__global__ void __launch_bounds__(256) peak_ffma_kernel(float* __restrict__ out, float a, float b)
{
    float x[PEAK_ILP];
#pragma unroll
    for (int k = 0; k < PEAK_ILP; k++)
        x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < PEAK_ILP; k++)
                x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_ILP; k++)
        s += x[k];
    if (s == 123.456f) // never: keeps the chains alive
        out[0] = s;
}

__global__ void __launch_bounds__(256) peak_ffma2_kernel(float* __restrict__ out, float a, float b)
{
    unsigned long long x[PEAK_ILP], A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int k = 0; k < PEAK_ILP; k++) {
        const float v = (float)(threadIdx.x + k) * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(v), "f"(v + 1.f));
    }
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < PEAK_ILP; k++)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(A), "l"(B));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_ILP; k++)
        s ^= x[k];
    if (s == 0x123456789abcdefull)
        out[0] = 1.f;
}

template <class K>
int time_kernel(aqc_ctx* ctx, K kern, unsigned grid, float* scratch, double flop_per_launch, double* tflops)
{
    cudaEvent_t e0, e1;
    AQC_CUDA(ctx, cudaEventCreate(&e0));
    AQC_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) { // the first repetition is the warm-up
        AQC_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        kern<<<grid, 256, 0, ctx->stream>>>(scratch, 0.999f, 1e-4f);
        AQC_LAUNCH_CHECK(ctx);
        AQC_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        AQC_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        AQC_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms > 0.f && flop_per_launch / (ms * 1e-3) > best)
            best = flop_per_launch / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best * 1e-12;
    return AQC_OK;
}

} // namespace

extern "C" int aqc_fp32_peak(aqc_ctx* ctx, double* tflops_ffma, double* tflops_ffma2)
{
    if (!ctx)
        return AQC_ERR_ARG;
    float* scratch = nullptr;
    AQC_CUDA(ctx, cudaMalloc(&scratch, 256));
    const unsigned grid = (unsigned)ctx->sm_count * 8u; // 2048 threads per SM: every scheduler full
    const double fmas = (double)grid * 256.0 * PEAK_ILP * 4.0 * PEAK_ITERS;
    double a = 0.0, b = 0.0;
    int rc = time_kernel(ctx, peak_ffma_kernel, grid, scratch, 2.0 * fmas, &a);
    if (!rc)
        rc = time_kernel(ctx, peak_ffma2_kernel, grid, scratch, 4.0 * fmas, &b);
    cudaFree(scratch);
    if (tflops_ffma)
        *tflops_ffma = a;
    if (tflops_ffma2)
        *tflops_ffma2 = b;
    return rc;
}
