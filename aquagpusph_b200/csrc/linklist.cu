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
//   prepare  : cell keys + the digit totals of every sort pass + ihoc = N, one pass
//   sort     : LSD radix, digits of 8 - 11 bits (the reference uses 4): TWO passes
//              for the 18 - 22 bit cell keys of the BASELINE cases, one kernel per
//              pass (tile offsets by decoupled look-back instead of a histogram,
//              a scan and a scatter launch); tiles of 4096 keys per CTA, ranking
//              with __match_any_sync (stable inside a warp by lane order, across
//              warps/CTAs by prefix order), keys and permutation staged through
//              shared memory so global stores of a digit run are contiguous; the
//              last pass also emits the inverse permutation
//              (RadixSort.cl.in:313-323).
//   heads    : one pass over the sorted keys.
// 5 launches per build (min/max, prepare, 2 passes, heads) against 12 before.
#include <math.h>

#include "aqc_common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 4096 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;

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
struct Pos3 { float x, y, z; };
template <int VS>
__device__ __forceinline__ Pos3 load_pos(const float* __restrict__ r, size_t i)
{
    Pos3 q;
    if constexpr (VS == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r) + i);
        q.x = t.x; q.y = t.y; q.z = t.z;
    } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(r) + i);
        q.x = t.x; q.y = t.y; q.z = 0.f;
    }
    return q;
}
template <int VS>
__device__ __forceinline__ uint32_t icell_of(const Pos3 q, float rminx, float rminy, float rminz,
                                             float idist, uint32_t nx, uint32_t ny)
{
    // explicit _rn intrinsics: never contracted, IEEE like the reference
    const uint32_t cx = (uint32_t)__fmul_rn(__fsub_rn(q.x, rminx), idist) + 3u;
    const uint32_t cy = (uint32_t)__fmul_rn(__fsub_rn(q.y, rminy), idist) + 3u;
    uint32_t id = cx - 1u + (cy - 1u) * nx;
    if constexpr (VS == 4) {
        const uint32_t cz = (uint32_t)__fmul_rn(__fsub_rn(q.z, rminz), idist) + 3u;
        id += (cz - 1u) * nx * ny;
    }
    return id;
}

// ---- radix sort ---------------------------------------------------------------
// Stable LSD sort, ONE kernel per digit (decoupled look-back between the tiles of a pass instead of
// a histogram + scan + scatter triple) and digits of 8 ... 11 bits chosen so that the keys of a
// link-list (18 - 22 bits at the BASELINE sizes) need two passes.  Scratch (ctx->sort_hist):
//   ghist[MAX_PASSES][MAX_RADIX]  digit totals of every pass, counted by the prepare kernel
//   ticket[MAX_PASSES]            tile tickets (a tile's predecessors are always running or done)
//   status[pass][tile][RADIX]     tile digit counts: value | AGGREGATE, later prefix | INCLUSIVE
constexpr int MAX_PASSES = 4;
constexpr int MAX_BITS = 11;
constexpr int MAX_RADIX = 1 << MAX_BITS;
constexpr uint32_t ST_AGG = 1u << 30, ST_INC = 1u << 31, ST_VAL = ST_AGG - 1u;
constexpr int GH_WORDS = MAX_PASSES * MAX_RADIX + 32; // ghist + tickets (padded)

struct SortPlan {
    int passes;
    int bits[MAX_PASSES];
    int shift[MAX_PASSES];
};

// Prepare kernel: (optionally) the cell keys of the particles, the digit totals of every pass, the
// status words zeroed, and (optionally) ihoc filled with N (LinkList.cl.in:32-42) -- everything the
// passes and the heads kernel need that does not depend on the order of the keys.
template <int VS> // 0: keys are given, 2 / 4: keys = icell(r)
__global__ void __launch_bounds__(256)
sort_prepare_kernel(uint32_t* __restrict__ keys, const float* __restrict__ r, uint32_t n,
                    float rminx, float rminy, float rminz, float idist, uint32_t nx, uint32_t ny,
                    SortPlan plan, uint32_t* __restrict__ ghist, uint32_t* __restrict__ status,
                    size_t status_words, uint32_t* __restrict__ fill, uint32_t fill_n,
                    uint32_t fill_value, uint32_t* __restrict__ minmax_reset)
{
    extern __shared__ uint32_t sh[]; // sum over passes of 2^bits counters
    int total_bins = 0;
    for (int p = 0; p < plan.passes; p++)
        total_bins += 1 << plan.bits[p];
    for (int k = threadIdx.x; k < total_bins; k += blockDim.x)
        sh[k] = 0;
    __syncthreads();
    const int l = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    // a warp takes 4 x 32 consecutive particles per iteration: the four loads of a lane are in
    // flight together, and whole warps iterate together (__match_any_sync needs every lane)
    constexpr int U = 4;
    const size_t nwarps = stride >> 5;
    for (size_t base = (first >> 5) * (32 * U); base < n; base += nwarps * (32 * U)) {
        uint32_t key[U];
        bool valid[U];
        if constexpr (VS == 0) {
#pragma unroll
            for (int k = 0; k < U; k++) {
                const size_t i = base + k * 32 + l;
                valid[k] = i < n;
                key[k] = valid[k] ? __ldg(keys + i) : 0u;
            }
        } else {
            Pos3 q[U];
#pragma unroll
            for (int k = 0; k < U; k++) {
                const size_t i = base + k * 32 + l;
                valid[k] = i < n;
                if (valid[k])
                    q[k] = load_pos<VS>(r, i);
            }
#pragma unroll
            for (int k = 0; k < U; k++)
                if (valid[k]) {
                    key[k] = icell_of<VS>(q[k], rminx, rminy, rminz, idist, nx, ny);
                    keys[base + k * 32 + l] = key[k];
                }
        }
        // (passes unrolled to the maximum: plan.* is then read at constant offsets of the parameter
        // space instead of through a local copy)
#pragma unroll
        for (int p = 0; p < MAX_PASSES; p++) {
            if (p < plan.passes) {
                int off = 0;
#pragma unroll
                for (int q = 0; q < MAX_PASSES; q++)
                    if (q < p)
                        off += 1 << plan.bits[q];
                const uint32_t mask = (1u << plan.bits[p]) - 1u;
                const int shift = plan.shift[p];
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const uint32_t d = valid[k] ? ((key[k] >> shift) & mask) : 0xFFFFFFFFu;
                    const uint32_t m = __match_any_sync(0xffffffffu, d);
                    if (valid[k] && l == (__ffs(m) - 1))
                        atomicAdd(&sh[off + d], __popc(m));
                }
            }
        }
    }
    for (size_t k = first; k < status_words; k += stride)
        status[k] = 0u;
    for (size_t k = first; k < fill_n; k += stride)
        fill[k] = fill_value;
    if (minmax_reset && blockIdx.x == 0 && threadIdx.x < 8)
        minmax_reset[threadIdx.x] = threadIdx.x < 4 ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    int off = 0;
    for (int p = 0; p < plan.passes; p++) {
        const int R = 1 << plan.bits[p];
        for (int k = threadIdx.x; k < R; k += blockDim.x) {
            const uint32_t c = sh[off + k];
            if (c)
                atomicAdd(ghist + p * MAX_RADIX + k, c);
        }
        off += R;
    }
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

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sum of the digit-d counts of the tiles before `tile`: every status word carries its own flag, so
// no fence is needed; LOOK_W predecessors are polled per round (their loads overlap), which bounds
// the walk when a whole wave of tiles starts at once.  (Walking the DPT digits of a thread together
// was measured and dropped: more registers in the ranking loop, passes 10 % slower.)
constexpr int LOOK_W = 8;
template <int RADIX>
__device__ __forceinline__ uint32_t look_back(const uint32_t* __restrict__ status, int tile, int d)
{
    uint32_t excl = 0, polls = 0;
    int t = tile - 1;
    while (t >= 0) {
        if (++polls > (1u << 24)) // watchdog: a lost status word must not hang the GPU
            __trap();
        uint32_t s[LOOK_W];
#pragma unroll
        for (int k = 0; k < LOOK_W; k++)
            s[k] = (t - k >= 0) ? ld_status(status + (size_t)(t - k) * RADIX + d) : ST_INC;
        int used = 0;
#pragma unroll
        for (int k = 0; k < LOOK_W; k++) {
            if (used == k) { // everything before was an aggregate
                if (s[k] & (ST_AGG | ST_INC)) {
                    excl += s[k] & ST_VAL;
                    used = (s[k] & ST_INC) ? LOOK_W + 1 : k + 1;
                }
            }
        }
        if (used > LOOK_W)
            return excl;
        t -= used; // used < LOOK_W: the next one is not published yet, poll again from there
    }
    return excl;
}

// One pass over one tile.  vals_in == nullptr => values are the global indices (first pass,
// RadixSort.cl.in:35-52 "init").
template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
sort_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                 uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                 uint32_t* __restrict__ inv_out, uint32_t n, int shift,
                 const uint32_t* __restrict__ ghist, uint32_t* __restrict__ ticket,
                 uint32_t* __restrict__ status)
{
    constexpr int RADIX = 1 << BITS;
    constexpr int DPT = RADIX / SORT_THREADS; // digits per thread: d = threadIdx.x + j * 256
    extern __shared__ uint32_t smem[];
    uint32_t* skeys = smem;                         // SORT_TILE
    uint32_t* svals = skeys + SORT_TILE;            // SORT_TILE
    uint32_t* goff = svals + SORT_TILE;             // RADIX: global offset - local start
    uint16_t* dstart = (uint16_t*)(goff + RADIX);   // RADIX: local start of every digit
    uint16_t* wcnt = dstart + RADIX;                // SORT_WARPS x RADIX
    __shared__ uint32_t ws[SORT_WARPS];
    __shared__ uint32_t tile_s;

    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0)
        tile_s = atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < SORT_WARPS * RADIX / 2; k += SORT_THREADS)
        ((uint32_t*)wcnt)[k] = 0u;
    __syncthreads();
    const uint32_t tile = tile_s;
    const uint32_t base = tile * SORT_TILE;
    const uint32_t valid_count = min((uint32_t)SORT_TILE, n - base);
    const uint32_t lt_mask = (1u << l) - 1u;
    uint16_t* myw = wcnt + w * RADIX;

    uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
    uint16_t rank[SORT_ITEMS];
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
            old = myw[d];
            myw[d] = (uint16_t)(old + __popc(m));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = (uint16_t)(old + __popc(m & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // thread t owns the digits t + 256 j: prefix over the warps, publish the tile's counts, scan
    // over the digits (local starts and global digit bases), then the look-back
    uint32_t cnt[DPT];
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = threadIdx.x + j * SORT_THREADS;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; k++) {
            const uint32_t c = wcnt[k * RADIX + d];
            wcnt[k * RADIX + d] = (uint16_t)run;
            run += c;
        }
        cnt[j] = run;
        st_status(status + (size_t)tile * RADIX + d, run | (tile == 0 ? ST_INC : ST_AGG));
    }
    uint32_t lcarry = 0, gcarry = 0;
    uint32_t gbase[DPT];
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = threadIdx.x + j * SORT_THREADS;
        uint32_t lt, gt;
        const uint32_t ls = block_excl_scan_256(cnt[j], ws, &lt);
        const uint32_t gs = block_excl_scan_256(__ldg(ghist + d), ws, &gt);
        dstart[d] = (uint16_t)(lcarry + ls);
        gbase[j] = gcarry + gs - (lcarry + ls);
        lcarry += lt;
        gcarry += gt;
    }
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = threadIdx.x + j * SORT_THREADS;
        uint32_t excl = 0;
        if (tile > 0) {
            excl = look_back<RADIX>(status, (int)tile, d);
            st_status(status + (size_t)tile * RADIX + d, (excl + cnt[j]) | ST_INC);
        }
        goff[d] = gbase[j] + excl;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & (RADIX - 1);
        const uint32_t lp = (uint32_t)dstart[d] + myw[d] + rank[k];
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

// ---- iHoc + linkList (LinkList.cl.in:32-42, 92-113) ------------------------
// (ihoc was filled with N by the prepare kernel; this one also leaves the digit totals and the
// tickets of the sort at zero for the next build)
__global__ void __launch_bounds__(256)
heads_kernel(const uint32_t* __restrict__ icell, uint32_t* __restrict__ ihoc, uint32_t N,
             uint32_t* __restrict__ ghist_reset)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (ghist_reset && t < GH_WORDS)
        ghist_reset[t] = 0u;
    // four consecutive keys per thread (one 16-byte load); the key before them comes from the
    // neighbour lane, lane 0 of a warp fetches it
    const uint32_t i0 = t * 4u;
    uint32_t c[4] = { 0u, 0u, 0u, 0u };
    if (i0 + 3u < N) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(icell) + t);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
        for (uint32_t k = 0; k < 4u; k++)
            if (i0 + k < N)
                c[k] = __ldg(icell + i0 + k);
    }
    uint32_t prev = __shfl_up_sync(0xffffffffu, c[3], 1);
    if ((threadIdx.x & 31) == 0 && i0 > 0 && i0 < N)
        prev = __ldg(icell + i0 - 1u);
    if (N < 2) // the reference launches linkList on N - 1 work-items: a single particle sets no head
        return;
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
        if (i0 + k < N && (i0 + k == 0 || prev != c[k]))
            ihoc[c[k]] = i0 + k;
        prev = c[k];
    }
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

SortPlan make_plan(uint32_t key_max)
{
    // digits covering every key < key_max (0 => full 32 bit): as few passes as 11-bit digits allow,
    // the bits spread evenly over them (never less than 8: a narrower digit is not cheaper)
    int bits = 32;
    if (key_max) {
        uint32_t top = key_max - 1;
        bits = 0;
        while (top) {
            bits++;
            top >>= 1;
        }
        if (bits == 0)
            bits = 1;
    }
    SortPlan pl{};
    pl.passes = (bits + MAX_BITS - 1) / MAX_BITS;
    int shift = 0;
    for (int p = 0; p < pl.passes; p++) {
        int b = (bits - shift + (pl.passes - p) - 1) / (pl.passes - p);
        if (b < 8)
            b = 8;
        pl.bits[p] = b;
        pl.shift[p] = shift;
        shift += b;
    }
    return pl;
}

size_t status_words(const SortPlan& pl, size_t nblocks)
{
    size_t w = 0;
    for (int p = 0; p < pl.passes; p++)
        w += nblocks << pl.bits[p];
    return w;
}

int ensure_sort_scratch(aqc_ctx* ctx, size_t n, const SortPlan& pl)
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
    const size_t hneed = GH_WORDS + status_words(pl, nblocks);
    if (hneed > ctx->sort_hist_cap) {
        if (ctx->sort_hist)
            AQC_CUDA(ctx, cudaFree(ctx->sort_hist));
        ctx->sort_hist = nullptr;
        ctx->sort_hist_cap = 0;
        ctx->sort_ghist_clean = false;
        AQC_CUDA(ctx, cudaMalloc(&ctx->sort_hist, (hneed + hneed / 8) * sizeof(uint32_t)));
        ctx->sort_hist_cap = hneed + hneed / 8;
    }
    return AQC_OK;
}

constexpr size_t pass_smem(int bits)
{
    return (size_t)2 * SORT_TILE * 4 + ((size_t)4 << bits) + ((size_t)2 << bits) +
           ((size_t)2 * SORT_WARPS << bits);
}

template <int BITS>
int launch_pass(aqc_ctx* ctx, uint32_t nblocks, const uint32_t* kin, const uint32_t* vin,
                uint32_t* kout, uint32_t* vout, uint32_t* inv, uint32_t n, int shift,
                const uint32_t* ghist, uint32_t* ticket, uint32_t* status)
{
    constexpr size_t smem = pass_smem(BITS);
    // (per device, and cheap: set before every launch instead of remembering it per process)
    AQC_CUDA(ctx, cudaFuncSetAttribute(sort_pass_kernel<BITS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_pass_kernel<BITS><<<nblocks, SORT_THREADS, smem, ctx->stream>>>(
        kin, vin, kout, vout, inv, n, shift, ghist, ticket, status);
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

// Digit totals and tickets must be zero when the prepare kernel starts.
int clean_ghist(aqc_ctx* ctx)
{
    if (!ctx->sort_ghist_clean)
        AQC_CUDA(ctx, cudaMemsetAsync(ctx->sort_hist, 0, GH_WORDS * sizeof(uint32_t), ctx->stream));
    ctx->sort_ghist_clean = false; // the prepare kernel is about to count into them
    return AQC_OK;
}

unsigned prepare_grid(aqc_ctx* ctx, size_t n)
{
    unsigned grid = aqc_blocks(n, 256 * 8);
    const unsigned cap = (unsigned)ctx->sm_count * 8;
    return grid > cap ? cap : (grid ? grid : 1);
}
size_t prepare_smem(const SortPlan& pl)
{
    size_t bins = 0;
    for (int p = 0; p < pl.passes; p++)
        bins += (size_t)1 << pl.bits[p];
    return bins * sizeof(uint32_t);
}

// Sort `n` keys (values implicit iota).  The prepare kernel has run: first_in holds the keys,
// digit totals counted, status words zero.  The sorted keys end in keys_out and the permutation in
// perm_out / inv_out (user arrays; perm_out / inv_out may be NULL).  Intermediate passes ping-pong
// in scratch, starting with sort_keys[first_out].
int run_sort(aqc_ctx* ctx, uint32_t n, const SortPlan& pl, const uint32_t* first_in, int first_out,
             uint32_t* keys_out, uint32_t* perm_out, uint32_t* inv_out)
{
    const uint32_t nblocks = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t* ghist = ctx->sort_hist;
    uint32_t* tickets = ctx->sort_hist + MAX_PASSES * MAX_RADIX;
    uint32_t* status = ctx->sort_hist + GH_WORDS;
    const uint32_t* kin = first_in;
    const uint32_t* vin = nullptr;
    int out = first_out;
    for (int p = 0; p < pl.passes; p++) {
        const bool last = (p == pl.passes - 1);
        uint32_t* kout = last ? keys_out : ctx->sort_keys[out];
        uint32_t* vout = last ? perm_out : ctx->sort_vals[out];
        uint32_t* inv = last ? inv_out : nullptr;
        const uint32_t* gh = ghist + p * MAX_RADIX;
        int rc;
        switch (pl.bits[p]) {
            case 8: rc = launch_pass<8>(ctx, nblocks, kin, vin, kout, vout, inv, n, pl.shift[p], gh, tickets + p, status); break;
            case 9: rc = launch_pass<9>(ctx, nblocks, kin, vin, kout, vout, inv, n, pl.shift[p], gh, tickets + p, status); break;
            case 10: rc = launch_pass<10>(ctx, nblocks, kin, vin, kout, vout, inv, n, pl.shift[p], gh, tickets + p, status); break;
            default: rc = launch_pass<11>(ctx, nblocks, kin, vin, kout, vout, inv, n, pl.shift[p], gh, tickets + p, status); break;
        }
        if (rc)
            return rc;
        status += (size_t)nblocks << pl.bits[p];
        kin = kout;
        vin = vout;
        out ^= 1;
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
    const SortPlan pl = make_plan(key_max);
    int rc = ensure_sort_scratch(ctx, n, pl);
    if (rc)
        return rc;
    if ((rc = clean_ghist(ctx)))
        return rc;
    // no pass may write the buffer it reads, and the last one writes `keys`: with one pass the
    // input is staged in scratch, with more the first pass reads `keys` where they lie.  Without a
    // permutation array the values still travel between the passes (scratch).
    const uint32_t* first_in = keys;
    if (pl.passes == 1) {
        AQC_CUDA(ctx, cudaMemcpyAsync(ctx->sort_keys[1], keys, (size_t)n * sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        first_in = ctx->sort_keys[1];
    }
    const uint32_t nblocks = (n + SORT_TILE - 1) / SORT_TILE;
    sort_prepare_kernel<0><<<prepare_grid(ctx, n), 256, prepare_smem(pl), ctx->stream>>>(
        const_cast<uint32_t*>(first_in), nullptr, n, 0.f, 0.f, 0.f, 0.f, 0u, 0u, pl, ctx->sort_hist,
        ctx->sort_hist + GH_WORDS, status_words(pl, nblocks), nullptr, 0u, 0u, nullptr);
    AQC_LAUNCH_CHECK(ctx);
    return run_sort(ctx, n, pl, first_in, 0, keys, perm, inv_perm);
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
        if (!ctx->minmax_clean) {
            minmax_init_kernel<<<1, 32, 0, ctx->stream>>>(ctx->minmax_dev);
            AQC_LAUNCH_CHECK(ctx);
        }
        ctx->minmax_clean = false;
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
    const SortPlan pl = make_plan(ncells[3]);
    int rc = ensure_sort_scratch(ctx, N, pl);
    if (rc)
        return rc;
    if ((rc = clean_ghist(ctx)))
        return rc;
    // one launch: cell keys (-> scratch), digit totals of every pass, status words zeroed, every
    // cell of ihoc = N, and the min/max slots back at their identities for the next build
    const float idist = 1.f / cell_length;
    const uint32_t nblocks = (N + SORT_TILE - 1) / SORT_TILE;
    uint32_t* keys0 = ctx->sort_keys[0];
    if (vs == 4)
        sort_prepare_kernel<4><<<prepare_grid(ctx, N), 256, prepare_smem(pl), ctx->stream>>>(
            keys0, (const float*)r, N, rmin[0], rmin[1], rmin[2], idist, ncells[0], ncells[1], pl,
            ctx->sort_hist, ctx->sort_hist + GH_WORDS, status_words(pl, nblocks), *ihoc, ncells[3],
            (uint32_t)N, ctx->minmax_dev);
    else
        sort_prepare_kernel<2><<<prepare_grid(ctx, N), 256, prepare_smem(pl), ctx->stream>>>(
            keys0, (const float*)r, N, rmin[0], rmin[1], 0.f, idist, ncells[0], ncells[1], pl,
            ctx->sort_hist, ctx->sort_hist + GH_WORDS, status_words(pl, nblocks), *ihoc, ncells[3],
            (uint32_t)N, ctx->minmax_dev);
    AQC_LAUNCH_CHECK(ctx);
    ctx->minmax_clean = true;
    rc = run_sort(ctx, N, pl, keys0, 1, icell, perm, inv_perm);
    if (rc)
        return rc;
    // iHoc: the heads (every other cell holds N since the prepare kernel)
    {
        const size_t threads = ((size_t)N + 3) / 4;
        heads_kernel<<<aqc_blocks(threads > GH_WORDS ? threads : GH_WORDS, 256), 256, 0, ctx->stream>>>(
            icell, *ihoc, N, ctx->sort_hist);
    }
    AQC_LAUNCH_CHECK(ctx);
    ctx->sort_ghist_clean = true;
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
