// linklist.cu -- LinkList / RadixSort / UnSort tools for sm_100a.
//
// Replaces aquagpusph/CalcServer/LinkList.cpp:326-494 (+ LinkList.cl.in),
// RadixSort.cpp:129-303 (+ RadixSort.cl.in) and UnSort.cl.in:30-42.
// Outputs are bit-exact with the reference: same r_min (min/max are order
// independent), IEEE division for idist, truncating float->uint conversion,
// stable sort, ihoc = first sorted index of every cell (N when empty).
//
// B200 design: all stages are HBM-bound integer work.
//   minmax   : one grid-stride pass, warp-shuffle + ordered-uint atomics
//   icell    : one pass, 16 B/particle read (3-D), 4 B write
//   sort     : LSD radix, 8 bit digits (the reference uses 4), tiles of 4096
//              keys per CTA; ranking with __match_any_sync (stable inside a
//              warp by lane order, across warps/CTAs by prefix order), keys and
//              permutation staged through shared memory so global stores of a
//              digit run are contiguous; the last pass also emits the inverse
//              permutation (RadixSort.cl.in:313-323).
//   heads    : one pass over the sorted keys.
#include <math.h>

#include "aqc_common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 4096 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float ord2f_host(uint32_t k)
{
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// ---- min / max of the positions (LinkList.cpp:80-91) ----------------------
template <int VS>
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ r, uint32_t N, uint32_t* __restrict__ out)
{
    float mn[VS], mx[VS];
#pragma unroll
    for (int c = 0; c < VS; c++) {
        mn[c] = INFINITY;
        mx[c] = -INFINITY;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N;
         i += (size_t)gridDim.x * blockDim.x) {
        float v[VS];
        if constexpr (VS == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(r) + i);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            const float2 t = __ldg(reinterpret_cast<const float2*>(r) + i);
            v[0] = t.x; v[1] = t.y;
        }
#pragma unroll
        for (int c = 0; c < VS; c++) {
            mn[c] = fminf(mn[c], v[c]);
            mx[c] = fmaxf(mx[c], v[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < VS; c++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    __shared__ float smn[8][VS], smx[8][VS];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
#pragma unroll
        for (int c = 0; c < VS; c++) {
            smn[w][c] = mn[c];
            smx[w][c] = mx[c];
        }
    __syncthreads();
    if (threadIdx.x < VS) {
        const int c = threadIdx.x;
        float a = smn[0][c], b = smx[0][c];
        for (int k = 1; k < 8; k++) {
            a = fminf(a, smn[k][c]);
            b = fmaxf(b, smx[k][c]);
        }
        atomicMin(out + c, f2ord(a));
        atomicMax(out + 4 + c, f2ord(b));
    }
}

__global__ void minmax_init_kernel(uint32_t* out)
{
    if (threadIdx.x < 4)
        out[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 8)
        out[threadIdx.x] = 0u;
}

// ---- iCell (LinkList.cl.in:54-85) -------------------------------------------
template <int VS>
__global__ void __launch_bounds__(256)
icell_kernel(uint32_t* __restrict__ icell, const float* __restrict__ r, uint32_t N,
             float rminx, float rminy, float rminz, float idist, uint32_t nx, uint32_t ny)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    float x, y, z = 0.f;
    if constexpr (VS == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r) + i);
        x = t.x; y = t.y; z = t.z;
    } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(r) + i);
        x = t.x; y = t.y;
    }
    // explicit _rn intrinsics: never contracted, IEEE like the reference
    const uint32_t cx = (uint32_t)__fmul_rn(__fsub_rn(x, rminx), idist) + 3u;
    const uint32_t cy = (uint32_t)__fmul_rn(__fsub_rn(y, rminy), idist) + 3u;
    uint32_t id = cx - 1u + (cy - 1u) * nx;
    if constexpr (VS == 4) {
        const uint32_t cz = (uint32_t)__fmul_rn(__fsub_rn(z, rminz), idist) + 3u;
        id += (cz - 1u) * nx * ny;
    }
    icell[i] = id;
}

// ---- radix sort ---------------------------------------------------------------
// Per-CTA digit histogram of one tile; hist layout [digit][block].
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                 uint32_t* __restrict__ hist, uint32_t nblocks)
{
    __shared__ uint32_t cnt[RADIX];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t idx = base + w * (32 * SORT_ITEMS) + k * 32 + l;
        const bool valid = idx < n;
        const uint32_t d = valid ? ((__ldg(keys + idx) >> shift) & (RADIX - 1)) : 0xFFFFFFFFu;
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (valid && l == (__ffs(m) - 1))
            atomicAdd(&cnt[d], __popc(m));
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* warp_sums,
                                                        uint32_t* total)
{
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (l >= o)
            x += y;
    }
    if (l == 31)
        warp_sums[w] = x;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SORT_WARPS; k++) {
        const uint32_t s = warp_sums[k];
        if (k < w)
            pre += s;
        tot += s;
    }
    if (total)
        *total = tot;
    __syncthreads();
    return pre + x - v;
}

// One CTA per digit: exclusive scan of hist[d][0..nblocks) in place, total -> tot[d]
__global__ void __launch_bounds__(SORT_THREADS)
sort_rowscan_kernel(uint32_t* __restrict__ hist, uint32_t nblocks, uint32_t* __restrict__ tot)
{
    __shared__ uint32_t ws[SORT_WARPS];
    uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nblocks; b0 += SORT_THREADS) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < nblocks ? row[b] : 0u;
        uint32_t t;
        const uint32_t e = block_excl_scan_256(v, ws, &t);
        if (b < nblocks)
            row[b] = carry + e;
        carry += t;
    }
    if (threadIdx.x == 0)
        tot[blockIdx.x] = carry;
}

// Stable scatter of one tile.  vals_in == nullptr => values are the global
// indices (first pass, RadixSort.cl.in:35-52 "init").
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    uint32_t* __restrict__ inv_out, uint32_t n, int shift,
                    const uint32_t* __restrict__ hist, uint32_t nblocks,
                    const uint32_t* __restrict__ tot)
{
    __shared__ uint32_t wcnt[SORT_WARPS][RADIX]; // 8 KB
    __shared__ uint32_t dstart[RADIX];           // local start of every digit
    __shared__ uint32_t goff[RADIX];             // global offset - local start
    __shared__ uint32_t ws[SORT_WARPS];
    __shared__ uint32_t skeys[SORT_TILE];        // 16 KB
    __shared__ uint32_t svals[SORT_TILE];        // 16 KB

    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const uint32_t base = blockIdx.x * SORT_TILE;
    const uint32_t valid_count = min((uint32_t)SORT_TILE, n - base);
    const uint32_t lt_mask = (1u << l) - 1u;

#pragma unroll
    for (int k = 0; k < SORT_WARPS; k++)
        wcnt[k][threadIdx.x] = 0;
    __syncthreads();

    uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t idx = base + w * (32 * SORT_ITEMS) + k * 32 + l;
        const bool valid = idx < n;
        key[k] = valid ? __ldg(keys_in + idx) : 0xFFFFFFFFu;
        val[k] = valid ? (vals_in ? __ldg(vals_in + idx) : idx) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & (RADIX - 1);
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (l == leader) {
            old = wcnt[w][d];
            wcnt[w][d] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(m & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // thread t owns digit t: prefix over warps, then scan over digits
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; k++) {
            const uint32_t c = wcnt[k][d];
            wcnt[k][d] = run;
            run += c;
        }
        const uint32_t ls = block_excl_scan_256(run, ws, nullptr);
        const uint32_t gb = block_excl_scan_256(__ldg(tot + d), ws, nullptr);
        dstart[d] = ls;
        goff[d] = gb + __ldg(hist + (size_t)d * nblocks + blockIdx.x) - ls;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & (RADIX - 1);
        const uint32_t lp = dstart[d] + wcnt[w][d] + rank[k];
        skeys[lp] = key[k];
        svals[lp] = val[k];
    }
    __syncthreads();

#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t lp = threadIdx.x + k * SORT_THREADS;
        if (lp < valid_count) {
            const uint32_t kk = skeys[lp];
            const uint32_t vv = svals[lp];
            const uint32_t pos = goff[(kk >> shift) & (RADIX - 1)] + lp;
            keys_out[pos] = kk;
            if (vals_out)
                vals_out[pos] = vv;
            if (inv_out)
                inv_out[vv] = pos;
        }
    }
}

__global__ void iota_kernel(uint32_t* p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = i;
}

// ---- iHoc + linkList (LinkList.cl.in:32-42, 92-113) ------------------------
__global__ void __launch_bounds__(256)
heads_kernel(const uint32_t* __restrict__ icell, uint32_t* __restrict__ ihoc, uint32_t N)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t c = __ldg(icell + i);
    if (i == 0 || __ldg(icell + i - 1) != c)
        ihoc[c] = i;
}

constexpr int MAX_FIELDS = 16;
struct ScatterParams {
    const void* src[MAX_FIELDS];
    void* dst[MAX_FIELDS];
    int bytes[MAX_FIELDS];
    int nfields;
};

__global__ void __launch_bounds__(256)
scatter_fields_kernel(const uint32_t* __restrict__ idx, uint32_t N, ScatterParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t o = __ldg(idx + i);
    for (int f = 0; f < P.nfields; f++) {
        switch (P.bytes[f]) {
            case 4:
                ((uint32_t*)P.dst[f])[o] = __ldg((const uint32_t*)P.src[f] + i);
                break;
            case 8:
                ((uint2*)P.dst[f])[o] = __ldg((const uint2*)P.src[f] + i);
                break;
            case 16:
                ((uint4*)P.dst[f])[o] = __ldg((const uint4*)P.src[f] + i);
                break;
            default: { // 64
                const uint4* s = (const uint4*)P.src[f] + (size_t)i * 4;
                uint4* d = (uint4*)P.dst[f] + (size_t)o * 4;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    d[k] = __ldg(s + k);
            }
        }
    }
}

int ensure_sort_scratch(aqc_ctx* ctx, size_t n)
{
    if (n > ctx->sort_cap) {
        const size_t cap = n + n / 8 + 1024;
        for (int k = 0; k < 2; k++) {
            if (ctx->sort_keys[k])
                AQC_CUDA(ctx, cudaFree(ctx->sort_keys[k]));
            if (ctx->sort_vals[k])
                AQC_CUDA(ctx, cudaFree(ctx->sort_vals[k]));
            ctx->sort_keys[k] = ctx->sort_vals[k] = nullptr;
        }
        ctx->sort_cap = 0;
        for (int k = 0; k < 2; k++) {
            AQC_CUDA(ctx, cudaMalloc(&ctx->sort_keys[k], cap * sizeof(uint32_t)));
            AQC_CUDA(ctx, cudaMalloc(&ctx->sort_vals[k], cap * sizeof(uint32_t)));
        }
        ctx->sort_cap = cap;
    }
    const size_t nblocks = (n + SORT_TILE - 1) / SORT_TILE;
    const size_t hneed = (size_t)RADIX * nblocks + RADIX;
    if (hneed > ctx->sort_hist_cap) {
        if (ctx->sort_hist)
            AQC_CUDA(ctx, cudaFree(ctx->sort_hist));
        ctx->sort_hist = nullptr;
        ctx->sort_hist_cap = 0;
        AQC_CUDA(ctx, cudaMalloc(&ctx->sort_hist, (hneed + hneed / 8) * sizeof(uint32_t)));
        ctx->sort_hist_cap = hneed + hneed / 8;
    }
    return AQC_OK;
}

int key_passes(uint32_t key_max)
{
    // number of 8-bit passes covering every key < key_max (0 => full 32 bit)
    if (key_max == 0)
        return 4;
    uint32_t top = key_max - 1;
    int bits = 0;
    while (top) {
        bits++;
        top >>= 1;
    }
    if (bits == 0)
        bits = 1;
    return (bits + RADIX_BITS - 1) / RADIX_BITS;
}

// Sort `n` keys found in ctx->sort_keys[start] (values implicit iota): the sorted
// keys end in keys_out and the permutation in perm_out / inv_out (user arrays;
// perm_out / inv_out may be NULL).  Intermediate passes ping-pong in scratch.

int run_sort(aqc_ctx* ctx, uint32_t n, int passes, int start, uint32_t* keys_out,
             uint32_t* perm_out, uint32_t* inv_out)
{
    const uint32_t nblocks = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t* hist = ctx->sort_hist;
    uint32_t* tot = ctx->sort_hist + (size_t)RADIX * nblocks;
    int cur = start;
    for (int p = 0; p < passes; p++) {
        const bool last = (p == passes - 1);
        const uint32_t* kin = ctx->sort_keys[cur];
        const uint32_t* vin = (p == 0) ? nullptr : ctx->sort_vals[cur];
        uint32_t* kout = last ? keys_out : ctx->sort_keys[cur ^ 1];
        uint32_t* vout = last ? perm_out : ctx->sort_vals[cur ^ 1];
        const int shift = p * RADIX_BITS;
        sort_hist_kernel<<<nblocks, SORT_THREADS, 0, ctx->stream>>>(kin, n, shift, hist, nblocks);
        AQC_LAUNCH_CHECK(ctx);
        sort_rowscan_kernel<<<RADIX, SORT_THREADS, 0, ctx->stream>>>(hist, nblocks, tot);
        AQC_LAUNCH_CHECK(ctx);
        sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, ctx->stream>>>(
            kin, vin, kout, vout, last ? inv_out : nullptr, n, shift, hist, nblocks, tot);
        AQC_LAUNCH_CHECK(ctx);
        cur ^= 1;
    }
    return AQC_OK;
}

} // namespace

extern "C" int aqc_radix_sort(aqc_ctx* ctx, aqc_usize* keys, aqc_usize n, aqc_usize key_max,
                              aqc_usize* perm, aqc_usize* inv_perm)
{
    if (!ctx || (!keys && n))
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_radix_sort: NULL keys");
    if (!n)
        return AQC_OK;
    aqc_pc_touch(ctx, keys, (size_t)n * sizeof(uint32_t));
    aqc_pc_touch(ctx, perm, (size_t)n * sizeof(uint32_t));
    aqc_pc_touch(ctx, inv_perm, (size_t)n * sizeof(uint32_t));
    int rc = ensure_sort_scratch(ctx, n);
    if (rc)
        return rc;
    const int passes = key_passes(key_max);
    // the last pass must not write the buffer it reads: stage the input in scratch
    const int start = 0;
    AQC_CUDA(ctx, cudaMemcpyAsync(ctx->sort_keys[start], keys, (size_t)n * sizeof(uint32_t),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    return run_sort(ctx, n, passes, start, keys, perm, inv_perm);
}

extern "C" int aqc_linklist_build(aqc_ctx* ctx, const void* r, aqc_usize N, int dims,
                                  float support, float h, int recompute_grid, float rmin[4],
                                  float rmax[4], aqc_usize ncells[4], aqc_usize* icell,
                                  aqc_usize** ihoc, size_t* ihoc_capacity, aqc_usize* perm,
                                  aqc_usize* inv_perm)
{
    if (!ctx || !r || !rmin || !rmax || !ncells || !icell || !ihoc || !ihoc_capacity || !perm ||
        !inv_perm)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_linklist_build: NULL argument");
    if (dims != 2 && dims != 3)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_linklist_build: dims must be 2 or 3");
    if (!N)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_linklist_build: N == 0");
    // LinkList.cpp:153-156,193-198: zero cell length is a setup error
    const float cell_length = support * h;
    if (!cell_length)
        return aqc_fail(ctx, AQC_ERR_ARG, "Zero cell length detected (Invalid number of cells)");
    // icell, ihoc and the permutations are rewritten (ihoc may also be replaced by a larger one)
    aqc_pc_touch(ctx, icell, (size_t)N * sizeof(aqc_usize));
    aqc_pc_touch(ctx, perm, (size_t)N * sizeof(aqc_usize));
    aqc_pc_touch(ctx, inv_perm, (size_t)N * sizeof(aqc_usize));
    aqc_pc_touch(ctx, *ihoc, (*ihoc_capacity ? *ihoc_capacity : 1) * sizeof(aqc_usize));

    const int vs = (dims == 3) ? 4 : 2;
    if (recompute_grid) {
        minmax_init_kernel<<<1, 32, 0, ctx->stream>>>(ctx->minmax_dev);
        AQC_LAUNCH_CHECK(ctx);
        unsigned grid = aqc_blocks(N, 256);
        const unsigned cap = (unsigned)ctx->sm_count * 8;
        if (grid > cap)
            grid = cap;
        if (vs == 4)
            minmax_kernel<4><<<grid, 256, 0, ctx->stream>>>((const float*)r, N, ctx->minmax_dev);
        else
            minmax_kernel<2><<<grid, 256, 0, ctx->stream>>>((const float*)r, N, ctx->minmax_dev);
        AQC_LAUNCH_CHECK(ctx);
        // multi-device: one global grid (addition to the reference, see mpi.cu)
        if (int rcc = aqc_comm_minmax(ctx, ctx->minmax_dev))
            return rcc;
        AQC_CUDA(ctx, cudaMemcpyAsync(ctx->minmax_host, ctx->minmax_dev, 8 * sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        // the reference blocks here too (LinkList.cpp:350-356)
        AQC_SYNC(ctx);
        const uint32_t* k = (const uint32_t*)ctx->minmax_host;
        for (int c = 0; c < 4; c++) {
            rmin[c] = (c < vs) ? ord2f_host(k[c]) : 0.f;
            rmax[c] = (c < vs) ? ord2f_host(k[4 + c]) : 0.f;
        }
        if (dims == 3) {
            // identities VEC_INFINITY / -VEC_INFINITY have w = 0 (Reduction.hcl.in:90)
            rmin[3] = fminf(0.f, rmin[3]);
            rmax[3] = fmaxf(-0.f, rmax[3]);
        }
    }
    // LinkList::nCells (LinkList.cpp:185-232), host fp32 like the reference
    {
        uint64_t n[3] = { 1, 1, 1 };
        for (int c = 0; c < dims; c++) {
            const volatile float span = rmax[c] - rmin[c];
            const volatile float q = span / cell_length;
            n[c] = (uint64_t)q + 6;
        }
        ncells[0] = (aqc_usize)n[0];
        ncells[1] = (aqc_usize)n[1];
        ncells[2] = (aqc_usize)n[2];
        const uint64_t w = n[0] * n[1] * n[2];
        if (w >= 0xFFFFFFFFull)
            return aqc_fail(ctx, AQC_ERR_ARG, "n_cells.w = %llu overflows 32-bit indices",
                            (unsigned long long)w);
        ncells[3] = (aqc_usize)w;
    }
    // LinkList::allocate (LinkList.cpp:234-271)
    if ((size_t)ncells[3] > *ihoc_capacity || !*ihoc) {
        if (*ihoc) {
            AQC_SYNC(ctx);
            AQC_CUDA(ctx, cudaFree(*ihoc));
            *ihoc = nullptr;
            *ihoc_capacity = 0;
        }
        void* p = nullptr;
        AQC_CUDA(ctx, cudaMalloc(&p, (size_t)ncells[3] * sizeof(aqc_usize)));
        *ihoc = (aqc_usize*)p;
        *ihoc_capacity = ncells[3];
    }
    int rc = ensure_sort_scratch(ctx, N);
    if (rc)
        return rc;
    const int passes = key_passes(ncells[3]);
    const int start = 0;
    const float idist = 1.f / cell_length;
    if (vs == 4)
        icell_kernel<4><<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>(
            ctx->sort_keys[start], (const float*)r, N, rmin[0], rmin[1], rmin[2], idist,
            ncells[0], ncells[1]);
    else
        icell_kernel<2><<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>(
            ctx->sort_keys[start], (const float*)r, N, rmin[0], rmin[1], 0.f, idist, ncells[0],
            ncells[1]);
    AQC_LAUNCH_CHECK(ctx);
    rc = run_sort(ctx, N, passes, start, icell, perm, inv_perm);
    if (rc)
        return rc;
    // iHoc: every cell = N, then the heads
    {
        const aqc_usize Nval = N;
        rc = aqc_fill(ctx, *ihoc, ncells[3], sizeof(aqc_usize), &Nval);
        if (rc)
            return rc;
    }
    if (N >= 2) { // the reference launches linkList on N-1 work-items
        heads_kernel<<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>(icell, *ihoc, N);
        AQC_LAUNCH_CHECK(ctx);
    }
    return AQC_OK;
}

extern "C" int aqc_scatter_fields(aqc_ctx* ctx, const aqc_usize* idx, aqc_usize N, int nfields,
                                  const void* const* src, void* const* dst,
                                  const size_t* elem_bytes)
{
    if (!ctx || !idx || !src || !dst || !elem_bytes)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_scatter_fields: NULL argument");
    if (!N || nfields <= 0)
        return AQC_OK;
    for (int f0 = 0; f0 < nfields; f0 += MAX_FIELDS) {
        ScatterParams P;
        P.nfields = (nfields - f0 < MAX_FIELDS) ? nfields - f0 : MAX_FIELDS;
        for (int f = 0; f < P.nfields; f++) {
            const size_t b = elem_bytes[f0 + f];
            if (b != 4 && b != 8 && b != 16 && b != 64)
                return aqc_fail(ctx, AQC_ERR_ARG, "aqc_scatter_fields: element size %zu", b);
            if (src[f0 + f] == dst[f0 + f])
                return aqc_fail(ctx, AQC_ERR_ARG, "aqc_scatter_fields: in-place field %d", f0 + f);
            P.src[f] = src[f0 + f];
            P.dst[f] = dst[f0 + f];
            aqc_pc_touch(ctx, dst[f0 + f], (size_t)N * b);
            P.bytes[f] = (int)b;
        }
        scatter_fields_kernel<<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>(idx, N, P);
        AQC_LAUNCH_CHECK(ctx);
    }
    return AQC_OK;
}
