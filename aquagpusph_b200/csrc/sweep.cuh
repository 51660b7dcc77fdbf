// sweep.cuh -- the neighbour sweep: the engines behind every kernel the reference
// writes with BEGIN_NEIGHS/END_NEIGHS (resources/Scripts/types/3D.h:197-219,
// 2D.h:174-193).
//
// Reference shape: one work-item per particle walks the 27 (9) cell lists on
// its own, fetching icell[j], imove[j], r[j], ... from global memory per
// candidate, and recomputing per-j factors (m_j/rho_j, kernel constants) for
// every pair.
//
// Here particles are cell-ordered (the link-list sort), so the particles of a
// cell are contiguous and a policy struct per script kernel (sweeps.cu) says what
// a pair needs: load_i, stage_j (position + hoisted per-j weights; excluded j get
// a far-away position), the exact test, the pair body, store_i.  Three engines:
//   * sweep_kernel   -- one warp per 32 particles, lanes grouped by cell, tiles of
//     32 candidates staged per warp, hits visited in the reference's order (cells
//     x-outer, ascending j): the order-dependent ElasticBounce / PST kernels;
//   * sweep2_kernel  -- the same walk with the packed-fp32 candidate filter:
//     kernels whose i particles are few and scattered (sensors, boundary elements,
//     halo sweeps) and small 2-D problems;
//   * sweep3_kernel  -- CTA-shared tiles in a shared-memory ring fed by a producer
//     warp, per-lane FIFOs of hit masks and deferred, lane-balanced pair bodies
//     (MODE 0); MODE 1 stores the hit masks of the filter (the pair-mask cache),
//     MODE 2 reads them instead of filtering, with its j rows packed by a pre-pass
//     and moved into the ring by TMA bulk copies.  DESIGN.md section 4 / 4.1.
#pragma once
#include "aqc_common.cuh"

struct LLParams {
    // cell of the i particles: the same array as `icell`, except for the remote
    // (halo) lists of cfd/MPI.cl, LINKLIST_REMOTE_PARAMS (types.h:117-122)
    const uint32_t* __restrict__ icell_i;
    const uint32_t* __restrict__ icell;
    const uint32_t* __restrict__ ihoc;
    uint32_t nx, ny, nz, nw; // n_cells (svec4)
    uint32_t N;
    // optional: classes of the particles of every cell (bits of aqc_cls_bit) and the classes
    // the kernel's j set is made of; a cell without any of them is not visited
    const uint8_t* __restrict__ cls = nullptr;
    uint32_t jmask = 0xFFu;
    // optional (remote lists): near[c] != 0 when one of the 3^D cells around c holds a j particle;
    // a warp none of whose particles sits in such a cell has nothing to do
    const uint8_t* __restrict__ near = nullptr;
};

__host__ __device__ inline uint32_t aqc_cls_bit(int mv)
{
    return mv == 1 ? 1u : mv == 0 ? 2u : mv == -1 ? 4u : mv == -2 ? 8u : mv == -3 ? 16u : 32u;
}

// cls[c] |= class bit of every particle of cell c (cls zeroed before; bytes or-ed through
// their 32-bit word, one atomic per run of equal (cell, class) inside a warp)
static __global__ void cell_class_kernel(const int* __restrict__ imove, const uint32_t* __restrict__ icell,
                                         uint32_t N, uint32_t nw, uint8_t* __restrict__ cls)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t c = __ldg(icell + i);
    if (c >= nw)
        return;
    const uint32_t bit = aqc_cls_bit(__ldg(imove + i));
    if (i > 0 && (threadIdx.x & 31) && __ldg(icell + i - 1) == c && aqc_cls_bit(__ldg(imove + i - 1)) == bit)
        return; // the previous lane does it
    atomicOr(reinterpret_cast<uint32_t*>(cls) + (c >> 2), bit << (8u * (c & 3u)));
}

// near[c'] = 1 for the 3^D cells c' around every cell that heads a run of the sorted list `icell`
// (near zeroed before).  The halo of a slab occupies a few layers of cells next to its cuts: the
// remote sweeps of cfd/MPI.cl visit every local particle, and all but those layers find 27 empty
// cells -- with the flags they leave after one byte load.
static __global__ void remote_near_kernel(const uint32_t* __restrict__ icell, uint32_t n, uint32_t nx, uint32_t ny,
                                          uint32_t nw, int dims, uint8_t* __restrict__ near)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t c = __ldg(icell + i);
    if (c >= nw || (i > 0 && __ldg(icell + i - 1) == c))
        return;
    const int kz = dims == 3 ? 1 : 0;
    for (int dz = -kz; dz <= kz; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const uint32_t cc = c + (uint32_t)dx + (uint32_t)dy * nx + (uint32_t)dz * nx * ny;
                if (cc < nw)
                    near[cc] = 1;
            }
}

constexpr float AQC_FAR = 3.0e38f; // staged position of an excluded j

constexpr int SWEEP_WARPS = 4;
constexpr int SWEEP_THREADS = SWEEP_WARPS * 32;

template <int DIMS>
__device__ __forceinline__ float dist2(float dx, float dy, float dz)
{
    if constexpr (DIMS == 3)
        return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    else
        return fmaf(dy, dy, dx * dx);
}

// Policy concept (see sweeps.cu):
//   static constexpr int DIMS, NJ4 (float4 slots staged per j)
//   struct IState
//   bool  i_active(int imove_i) const
//   void  load_i(IState&, uint32_t i) const
//   void  stage_j(uint32_t j, float4* out /*[NJ4]*/) const
//   bool  test(const IState&, const float4& A) const      -- candidate filter
//   void  body(IState&, const float4* row, int stride) const  -- row[k*stride]
//   void  store_i(const IState&, uint32_t i) const
template <class P, bool COMPACT>
__global__ void __launch_bounds__(SWEEP_THREADS)
sweep_kernel(const P p, const LLParams ll)
{
    __shared__ float4 sj[SWEEP_WARPS][P::NJ4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * SWEEP_WARPS + warp) * 32u + lane;
    const bool valid = i < ll.N;
    const bool active = valid && p.i_active(p.imove[valid ? i : 0]);
    uint32_t remaining = __ballot_sync(0xffffffffu, active);
    if (!remaining)
        return;
    const uint32_t c_i = active ? __ldg(ll.icell_i + i) : 0xFFFFFFFFu;
    if (ll.near && !__any_sync(0xffffffffu, active && c_i < ll.nw && ll.near[c_i < ll.nw ? c_i : 0]))
        return;
    typename P::IState st;
    if (active)
        p.load_i(st, i);
    float4(*tile)[32] = sj[warp];
    constexpr int KZ = (P::DIMS == 3) ? 1 : 0;

    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t c = __shfl_sync(0xffffffffu, c_i, leader);
        const bool mine = active && (c_i == c);
        remaining &= ~__ballot_sync(0xffffffffu, mine);

        for (int ci = -1; ci <= 1; ci++)
            for (int cj = -1; cj <= 1; cj++)
                for (int ck = -KZ; ck <= KZ; ck++) {
                    const uint32_t cell = c + (uint32_t)ci + (uint32_t)cj * ll.nx +
                                          (uint32_t)ck * ll.nx * ll.ny;
                    if (ll.cls && !(ll.cls[cell] & ll.jmask))
                        continue;
                    uint32_t j0 = __ldg(ll.ihoc + cell);
                    while (j0 < ll.N) {
                        const uint32_t jj = j0 + lane;
                        const bool in = (jj < ll.N) && (__ldg(ll.icell + jj) == cell);
                        const int cnt = __popc(__ballot_sync(0xffffffffu, in));
                        if (!cnt)
                            break;
                        bool live = false;
                        if (in) {
                            float4 o[P::NJ4];
                            p.stage_j(jj, o);
                            live = P::j_live(o[0]);
#pragma unroll
                            for (int k = 0; k < P::NJ4; k++)
                                tile[k][lane] = o[k];
                        }
                        // only the candidates that can interact at all are tested (a tile of
                        // fluid particles costs a boundary kernel one ballot)
                        const uint32_t lm = __ballot_sync(0xffffffffu, live);
                        __syncwarp();
                        if (mine && lm) {
                            if constexpr (COMPACT) {
                                uint32_t hits = 0;
                                for (uint32_t mm = lm; mm; mm &= mm - 1) {
                                    const int k = __ffs(mm) - 1;
                                    if (p.test(st, tile[0][k]))
                                        hits |= 1u << k;
                                }
                                while (hits) {
                                    const int k = __ffs(hits) - 1;
                                    hits &= hits - 1;
                                    p.body(st, &tile[0][k], 32);
                                }
                            } else {
                                for (uint32_t mm = lm; mm; mm &= mm - 1) {
                                    const int k = __ffs(mm) - 1;
                                    if (p.test(st, tile[0][k]))
                                        p.body(st, &tile[0][k], 32);
                                }
                            }
                        }
                        __syncwarp();
                        if (cnt < 32)
                            break;
                        j0 += 32;
                    }
                }
    }
    if (active)
        p.store_i(st, i);
}

// ---------------------------------------------------------------------------
// v2 engine for policies whose candidate filter is the kernel-support sphere
// (P::SPHERE): the 32 staged candidates are tested two at a time with packed
// fp32 (FFMA2, sm_100) on the expanded form
//     |r_j - r_i|^2 - cut^2 = (|r_i|^2 - cut^2) + |r_j|^2 - 2 r_i . r_j
// in coordinates relative to a warp-local origin (so the cancellation error is
// ~1e-6 of cut^2), and the sign bit of each result is shifted into the lane's
// hit mask with one funnel shift: 1 LDS.128 + 2 packed FP + 1 SHF per
// candidate instead of 12 instructions.  The filter radius is inflated by 1e-5
// so it can only over-select; every selected pair is re-tested exactly
// (P::test, absolute coordinates) before its body runs, hence the set of pairs
// and their order are exactly those of the scalar engine above.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Test data of one tile: pair q = candidates (2q, 2q+1):
//   T[2q] = (x0, x1, y0, y1)   T[2q+1] = (z0, z1, n0, n1),  n = x^2 + y^2 + z^2
// Returns the hit mask with candidate k at bit (31 - k).
template <int NPAIR>
__device__ __forceinline__ uint32_t test_tile(const float4* __restrict__ T, unsigned long long X2,
                                              unsigned long long Y2, unsigned long long Z2,
                                              unsigned long long C2)
{
    uint32_t hits = 0;
#pragma unroll
    for (int q = 0; q < NPAIR; q++) {
        const float4 a = T[2 * q], b = T[2 * q + 1];
        unsigned long long t = ffma2(X2, pack2(a.x, a.y), C2);
        t = ffma2(Y2, pack2(a.z, a.w), t);
        t = ffma2(Z2, pack2(b.x, b.y), t);
        t = fadd2(t, pack2(b.z, b.w));
        hits = __funnelshift_l((uint32_t)t, hits, 1);         // sign of candidate 2q
        hits = __funnelshift_l((uint32_t)(t >> 32), hits, 1); // sign of candidate 2q+1
    }
    return hits << (32 - 2 * NPAIR);
}

// The same test against two radii at once: `hits` for C2, `sure` for C2 + D2 (a smaller radius)
template <int NPAIR>
__device__ __forceinline__ void test_tile2(const float4* __restrict__ T, unsigned long long X2,
                                           unsigned long long Y2, unsigned long long Z2,
                                           unsigned long long C2, unsigned long long D2, uint32_t& hits,
                                           uint32_t& sure)
{
    uint32_t h = 0, s = 0;
#pragma unroll
    for (int q = 0; q < NPAIR; q++) {
        const float4 a = T[2 * q], b = T[2 * q + 1];
        unsigned long long t = ffma2(X2, pack2(a.x, a.y), C2);
        t = ffma2(Y2, pack2(a.z, a.w), t);
        t = ffma2(Z2, pack2(b.x, b.y), t);
        t = fadd2(t, pack2(b.z, b.w));
        const unsigned long long u = fadd2(t, D2);
        h = __funnelshift_l((uint32_t)t, h, 1);
        h = __funnelshift_l((uint32_t)(t >> 32), h, 1);
        s = __funnelshift_l((uint32_t)u, s, 1);
        s = __funnelshift_l((uint32_t)(u >> 32), s, 1);
    }
    hits = h << (32 - 2 * NPAIR);
    sure = s << (32 - 2 * NPAIR);
}

__device__ __forceinline__ float4 lds128(uint32_t a) // a: shared-window address
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(a)
                 : "memory");
    return v;
}

constexpr float AQC_NEVER = 1.0e30f; // |r_j|^2 of a padding / excluded candidate

template <class P>
__global__ void __launch_bounds__(SWEEP_THREADS)
sweep2_kernel(const P p, const LLParams ll)
{
    __shared__ float4 sT[SWEEP_WARPS][32];
    __shared__ float4 sj[SWEEP_WARPS][P::NJ4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * SWEEP_WARPS + warp) * 32u + lane;
    const bool valid = i < ll.N;
    const bool active = valid && p.i_active(p.imove[valid ? i : 0]);
    uint32_t remaining = __ballot_sync(0xffffffffu, active);
    if (!remaining)
        return;
    const uint32_t c_i = active ? __ldg(ll.icell_i + i) : 0xFFFFFFFFu;
    if (ll.near && !__any_sync(0xffffffffu, active && c_i < ll.nw && ll.near[c_i < ll.nw ? c_i : 0]))
        return;
    typename P::IState st;
    st.x = st.y = st.z = 0.f;
    if (active)
        p.load_i(st, i);
    float4(*tile)[32] = sj[warp];
    float* tst = reinterpret_cast<float*>(sT[warp]);
    const int tslot = (lane >> 1) * 8 + (lane & 1);
    constexpr int KZ = (P::DIMS == 3) ? 1 : 0;
    const float cut2f = p.cut2 * 1.00001f;

    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t c = __shfl_sync(0xffffffffu, c_i, leader);
        const bool mine = active && (c_i == c);
        remaining &= ~__ballot_sync(0xffffffffu, mine);
        // warp-local origin: the leader's position
        const float ox = __shfl_sync(0xffffffffu, st.x, leader);
        const float oy = __shfl_sync(0xffffffffu, st.y, leader);
        const float oz = (P::DIMS == 3) ? __shfl_sync(0xffffffffu, st.z, leader) : 0.f;
        const float xi = st.x - ox, yi = st.y - oy, zi = (P::DIMS == 3) ? st.z - oz : 0.f;
        const unsigned long long X2 = pack2(-2.f * xi, -2.f * xi);
        const unsigned long long Y2 = pack2(-2.f * yi, -2.f * yi);
        const unsigned long long Z2 = pack2(-2.f * zi, -2.f * zi);
        const float ci = fmaf(zi, zi, fmaf(yi, yi, xi * xi)) - cut2f;
        const unsigned long long C2 = pack2(ci, ci);

        for (int cx = -1; cx <= 1; cx++)
            for (int cy = -1; cy <= 1; cy++)
                for (int cz = -KZ; cz <= KZ; cz++) {
                    const uint32_t cell = c + (uint32_t)cx + (uint32_t)cy * ll.nx +
                                          (uint32_t)cz * ll.nx * ll.ny;
                    if (ll.cls && !(ll.cls[cell] & ll.jmask))
                        continue;
                    uint32_t j0 = __ldg(ll.ihoc + cell);
                    while (j0 < ll.N) {
                        const uint32_t jj = j0 + lane;
                        const bool in = (jj < ll.N) && (__ldg(ll.icell + jj) == cell);
                        const int cnt = __popc(__ballot_sync(0xffffffffu, in));
                        if (!cnt)
                            break;
                        float tx = 0.f, ty = 0.f, tz = 0.f, tn = AQC_NEVER;
                        if (in) {
                            float4 o[P::NJ4];
                            p.stage_j(jj, o);
#pragma unroll
                            for (int k = 0; k < P::NJ4; k++)
                                tile[k][lane] = o[k];
                            if (o[0].x != AQC_FAR) {
                                tx = o[0].x - ox;
                                ty = o[0].y - oy;
                                tz = (P::DIMS == 3) ? o[0].z - oz : 0.f;
                                tn = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
                            }
                        }
                        tst[tslot] = tx;
                        tst[tslot + 2] = ty;
                        tst[tslot + 4] = tz;
                        tst[tslot + 6] = tn;
                        __syncwarp();
                        if (mine) {
                            uint32_t hits;
                            if (cnt > 16)
                                hits = test_tile<16>(sT[warp], X2, Y2, Z2, C2);
                            else if (cnt > 8)
                                hits = test_tile<8>(sT[warp], X2, Y2, Z2, C2);
                            else
                                hits = test_tile<4>(sT[warp], X2, Y2, Z2, C2);
                            while (hits) {
                                const int k = __clz(hits);
                                hits ^= 0x80000000u >> k;
                                if (p.test(st, tile[0][k]))
                                    p.body(st, &tile[0][k], 32);
                            }
                        }
                        __syncwarp();
                        if (cnt < 32)
                            break;
                        j0 += 32;
                    }
                }
    }
    if (active)
        p.store_i(st, i);
}

// ---------------------------------------------------------------------------
// v3 engine (SPHERE policies): CTA-shared neighbour tiles, deferred pair bodies.
//
// v2 runs the pair bodies of a tile right after its filter, so a tile costs
// max-over-lanes(hits) body iterations: a lane's hit count in one neighbour cell
// depends on where its particle sits in its own cell, and the measured lane
// efficiency of the body loop is 0.40 (profiles/r1_ncu_shepard_v2a*); and every warp
// stages its own copy of every neighbour tile.  Here
//   * a CTA of 8 warps takes 256 consecutive (cell-ordered) particles; the particles
//     whose cells lie within S3_SPAN cells of the first one in its x row form a group
//     that is served by ONE walk over the union of the members' neighbourhoods: the
//     3 (2-D) or 9 (3-D) x rows [c0 - 1, c0 + span + 1] + offset, cut in parts of two
//     cells.  Tiles of 32 candidates are taken round-robin over the parts, so that the
//     tiles in flight always mix all directions; each warp stages one tile per round
//     (j rows + the packed test layout) into a shared ring of K rounds;
//   * a warp filters the tiles that hold cells adjacent to its lanes' cells and only
//     RECORDS the hit masks (per-warp mask ring); a lane keeps the candidates of ITS
//     OWN 3^D neighbourhood only (a tile holds at most two adjacent cells, the first
//     n1 candidates belong to the lower one), so the pair set is exactly the
//     reference's whatever the other lanes of the group need;
//   * every lane walks its masks with a FIFO cursor (seq, cur); a body iteration is
//     issued only while EVERY participating lane of the warp has a pending hit, or
//     when the oldest ring round has to be recycled.  Replaying this on the dam-break
//     state gives 0.80 (window of 8 tiles) to 0.92 (16+) lane efficiency.
// Per particle the hits are consumed in the order of a fixed traversal, so sums are
// deterministic (run-to-run bit-identical); the order is not the reference's x-outer
// one any more: results differ from v2 by fp32 rounding of the sums only.
// Order-dependent kernels stay on sweep_kernel.
// (a CTA of 9 warps -- 8 consumers at 56 registers -- is not faster: the mask-reading sweeps are
// bound by the issue slots and the shared-memory pipe of the SM, not by the number of warps)
constexpr int S3_WARPS = 8;                  // warp 0 stages tiles, warps 1..7 own 32 particles each
constexpr int S3_THREADS = S3_WARPS * 32;
constexpr int S3_CWARPS = S3_WARPS - 1;       // consumer warps
constexpr int S3_PARTICLES = S3_CWARPS * 32;  // particles of a CTA
#ifndef S3_TILES_N
#define S3_TILES_N 8
#endif
constexpr int S3_TILES = S3_TILES_N;           // tiles of a round
#ifndef S3_BATCH
#define S3_BATCH 1 // tiles whose global loads the producer keeps in flight together (divides S3_TILES)
#endif
#ifndef S3_PROFILE
#define S3_PROFILE 0 // debug builds only (tools/build_variant.py): per-phase clock64 totals
#endif
#if S3_PROFILE
static __device__ unsigned long long g_s3prof[32];
#define S3P_START long long s3p_t = clock64();
#define S3P_ACC(k)                                                                                 \
    {                                                                                              \
        const long long s3p_n = clock64();                                                         \
        if (lane == 0)                                                                             \
            atomicAdd(&g_s3prof[(k)], (unsigned long long)(s3p_n - s3p_t));                        \
        s3p_t = s3p_n;                                                                             \
    }
#else
#define S3P_START
#define S3P_ACC(k)
#endif
#ifndef S3_BBATCH
#define S3_BBATCH 4 // the same for the builders of the pair cache (MODE 1 / 3)
#endif
static_assert(S3_TILES_N % S3_BATCH == 0 && S3_TILES_N % S3_BBATCH == 0, "S3_BATCH must divide S3_TILES");
constexpr int S3_MAXK = 8;                    // ring rounds (at most; static shared memory grows with it)
#ifndef S3_RTILES
#define S3_RTILES 8 // tiles of a round of the mask-reading sweeps (MODE 2; divides S3_TILES)
#endif
#ifndef S3_RRING
#define S3_RRING 3 // their ring rounds (AQC_SWEEP_RING2)
#endif
static_assert(S3_TILES_N % S3_RTILES == 0, "S3_RTILES must divide S3_TILES");
static_assert(AQC_PC_ROUND_BYTES == (size_t)S3_TILES_N * (S3_WARPS - 1) * 32 * 4, "aqc_pairs_cache_stats");
constexpr int S3_SPAN = 7; // cells of one x row a group may span beyond the first
constexpr int S3_MAXE = 9 * ((S3_SPAN + 3 + 1) / 2);

// mbarrier (shared-memory arrive/wait barrier) used as the round barrier of the v3 engine: a
// warp that has arrived keeps consuming its pending hits while it polls, instead of idling
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t a)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t a, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
    return ok != 0;
}

// try_wait comes back after a short system-dependent time whatever the hint (measured: a third
// of the executed instructions were polls): a warp that finds its barrier closed sleeps
#ifndef S3_SLEEP_P
#define S3_SLEEP_P 400 // ns, producer warp (waits for ring space most of the time)
#endif
#ifndef S3_SLEEP_C
#define S3_SLEEP_C 100 // ns, consumer warps
#endif
#ifndef S3_WAIT_NS
#define S3_WAIT_NS 20000 // suspend-time hint of a waiting warp: it leaves the issue slots to the others
#endif
__device__ __forceinline__ bool mbar_wait(uint32_t a, uint32_t parity) // suspends the thread
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity), "r"((uint32_t)S3_WAIT_NS)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(a),
                 "r"(bytes)
                 : "memory");
}
// mbarrier.init must be visible to the async proxy (bulk copies complete on the barrier)
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// TMA 1-D bulk copy global -> shared (SASS UBLKCP): bytes % 16 == 0, both addresses 16-byte
// aligned; completes bytes of transaction count on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// first j in [b, N) whose cell (relative to lo) is beyond wid; icell is sorted
__device__ __forceinline__ uint32_t s3_run_end(const uint32_t* __restrict__ icell, uint32_t b,
                                               uint32_t N, uint32_t lo, uint32_t wid)
{
    uint32_t step = 32, good = b; // icell[good] is inside
    while (good + step < N && __ldg(icell + good + step) - lo <= wid) {
        good += step;
        step *= 2;
    }
    uint32_t bad = min(good + step, N); // first known outside (or N)
    while (bad - good > 1) {
        const uint32_t mid = good + (bad - good) / 2;
        if (__ldg(icell + mid) - lo <= wid)
            good = mid;
        else
            bad = mid;
    }
    return bad;
}

// Pair-mask cache of the v3 engine.  The hit masks of the candidate filter depend on the
// geometry only (positions, link-list, the classes of the particles): every sweep between two
// changes of those walks the same tiles and finds the same masks.  The reference's midpoint
// scheme keeps r fixed over the sub-iterations of a step (basic/time_scheme/midpoint.cl:93-111
// advances u and rho only), so MLS, the fluid pass and lapp_corr of all sub-iterations -- 7
// sweeps of the 3-D dam-break pipeline -- share one set.  MODE 1 (PMaskBuild) runs the filter
// alone and stores the masks, MODE 2 reads them instead of filtering; the exact test of every
// selected pair stays in the pair bodies, so the pair set and the summation order are those of
// MODE 0 (results bit-identical, tests/test_gpu_kernels.py).
//   masks    [round][S3_TILES][S3_CWARPS][32]   rounds allocated per (CTA, pass) by the builder
//   pass_tab [CTA][S3_MAXPASS]                   first round of the pass
//   ctl      [0] rounds allocated (keeps counting past the capacity: the host grows the buffer
//            and builds again), [1] bit 0: a CTA needed more than S3_MAXPASS passes
// the masks are written once and read once per sweep (1.2 KB per particle): streaming accesses
// (evict-first) keep them from pushing the packed rows and the particle arrays out of L2
#ifndef S3_MASK_CS
#define S3_MASK_CS 1
#endif
#if S3_MASK_CS
#define S3_MASK_LD(p) __ldcs(p)
#define S3_MASK_ST(p, v) __stcs(p, v)
#else
#define S3_MASK_LD(p) __ldg(p)
#define S3_MASK_ST(p, v) (*(p) = (v))
#endif
struct S3Cache {
    const float4* rows = nullptr; // MODE 2 / v4: the j rows of every particle, packed by s3_pack_kernel
    uint32_t* masks = nullptr;
    uint32_t* pass_tab = nullptr;
    unsigned long long* ctl = nullptr;
    uint32_t cap_rounds = 0;
    // neighbour lists (MODE 3 builds them, sweep4_kernel reads them; see "v4 engine" below)
    uint2* chunks = nullptr; // [CTA][consumer warp][capc][32 lanes]: 4 entries of 16 bits
    uint8_t* cnt = nullptr;  // [round][consumer warp][32 lanes]: chunks of the lane in that round
    uint32_t capc = 0;       // chunks a lane's list can hold
};
// An entry of a neighbour list is the byte offset of the candidate's row inside ONE row array of
// the reader's ring (S4_K rounds of S3_TILES tiles of 32 rows of 16 bytes); S4_NULL addresses the
// extra row behind the ring, whose weights are zero (the padding of a lane's last chunk in a round)
#ifndef S4_RING
#define S4_RING 5 // ring rounds of the list readers: how far the warps of a CTA may drift apart
#endif
constexpr int S4_K = S4_RING;
constexpr uint32_t S4_ROWS = S4_K * S3_TILES_N * 32;
constexpr uint32_t S4_NULL = S4_ROWS * 16;
constexpr int S3_MAXPASS = 32;
constexpr uint32_t S3_NOPASS = 0xFFFFFFFFu;

// MODE 2 pre-pass: the staged rows of EVERY particle (stage_j: position, hoisted per-j weights,
// exclusion) as NJ4 arrays of float4, so that the rows of a tile are NJ4 contiguous runs which
// the producer of sweep3_kernel moves with bulk copies instead of loading, converting and storing
template <class P>
__global__ void __launch_bounds__(256) s3_pack_kernel(const P p, const uint32_t N, float4* __restrict__ rows)
{
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= N)
        return;
    float4 o[P::NJ4];
    p.stage_j(j, o);
#pragma unroll
    for (int q = 0; q < P::NJ4; q++)
        rows[(size_t)q * N + j] = o[q];
}

template <class P, int MODE, int W>
__global__ void __launch_bounds__(S3_THREADS, (P::NJ4 <= 2) ? 4 : 3)
sweep3_kernel(const P p, const LLParams ll, const int K, const S3Cache pc)
{
    extern __shared__ float4 smem3[];
    constexpr int SLOT4 = P::NJ4 * 32;
    const uint32_t NS = (uint32_t)K * W; // ring slots
    float4* const sT = smem3;            // [K][W][32] packed test layout of the ring's tiles (not in MODE 2)
    float4* const sJ = sT + (MODE == 2 ? 0u : NS * 32); // [NS][SLOT4] j rows
    uint32_t NM = NS - W;                // FIFO entries per lane: the round being staged has no masks yet
    asm volatile("" : "+r"(NM));         // (opaque: kept in a register, not recomputed from K)
    uint32_t* const sM = reinterpret_cast<uint32_t*>(sJ + (size_t)NS * SLOT4); // [CW][NM][32] masks, then
                                                                               // [CW][NM][32] slot bytes
    __shared__ uint32_t e_begin[S3_MAXE], e_end[S3_MAXE], e_lo[S3_MAXE], e_rel[S3_MAXE];
    __shared__ uint32_t t_cnt[S3_MAXK][W], t_rel[S3_MAXK][W], t_n1[S3_MAXK][W];
    __shared__ uint32_t s_ball[S3_WARPS], s_c0, s_span, s_firstw, s_lastw, s_maxk, s_base;
    __shared__ float s_o[6];
    // full[k]: the producer has staged ring round k; empty[k]: a consumer warp is done with it
    __shared__ unsigned long long s_full[S3_MAXK], s_empty[S3_MAXK];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == 0;
    const int cw = warp - 1; // consumer index
    const uint32_t full_a = (uint32_t)__cvta_generic_to_shared(s_full);
    const uint32_t empty_a = (uint32_t)__cvta_generic_to_shared(s_empty);
    const uint32_t i = blockIdx.x * (uint32_t)S3_PARTICLES + (uint32_t)(tid - 32);
    const bool valid = !producer && i < ll.N;
    const bool active = valid && p.i_active(p.imove[valid ? i : 0]);
    // the passes of a CTA are made of ALL its particles, whatever the kernel's i set: the
    // tiles of a pass are then the same for every kernel (the mask cache relies on it)
    const uint32_t c_i = valid ? __ldg(ll.icell_i + i) : 0xFFFFFFFFu;
    if constexpr (P::REMOTE) {
        // remote (halo) list: a CTA none of whose particles has a halo particle in its 3^D cells has
        // nothing to do -- nor has the builder of its lists (same flags, same decision)
        if (ll.near && !__syncthreads_or(valid && c_i < ll.nw && ll.near[c_i < ll.nw ? c_i : 0]))
            return;
    }
    typename P::IState st;
    st.x = st.y = st.z = 0.f;
    if (active)
        p.load_i(st, i);
    constexpr int NROWS = (P::DIMS == 3) ? 9 : 3;
    const float cut2f = p.cut2 * 1.0001f;
    // shared-window addresses of this lane's FIFO (masks, slot bytes) and of the last row
    // slot of ring tile 0, kept opaque so that they stay in registers instead of being
    // recomputed from tid at every use
    uint32_t Mw_a = (uint32_t)__cvta_generic_to_shared(sM + (size_t)(producer ? 0 : cw) * NM * 32 + lane);
    uint32_t Sw_a = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<uint8_t*>(sM + (size_t)S3_CWARPS * NM * 32) +
                                                       (size_t)(producer ? 0 : cw) * NM * 32 + lane);
    uint32_t sJ_a = (uint32_t)__cvta_generic_to_shared(sJ) + 31 * 16;
    asm volatile("" : "+r"(Mw_a), "+r"(Sw_a), "+r"(sJ_a));

    // (kernels that never read the cache group their own i particles only: a boundary kernel
    // does not walk the neighbourhood of a CTA's fluid particles)
    uint32_t lc = 0; // MODE 3: chunks this lane has written to its list (a lane works in one pass)
    constexpr bool ALLPASS = P::CACHE || MODE == 1 || MODE == 3;
    bool pending = (ALLPASS ? valid : active) && c_i < ll.nw;
    for (uint32_t npass = 0;; npass++) {
        // ---- the group of this pass: first pending particle and its x-row neighbours
        const uint32_t pb = __ballot_sync(0xffffffffu, pending);
        if (lane == 0)
            s_ball[warp] = pb;
        if (tid == 0) {
            s_span = 0;
            s_firstw = 0xFFFFFFFFu;
            s_lastw = 0;
            s_maxk = 0;
            for (int k = 0; k < K; k++) { // nobody is inside a round loop here
                if (npass) {
                    mbar_inval(full_a + 8 * k);
                    mbar_inval(empty_a + 8 * k);
                }
                mbar_init(full_a + 8 * k, 1);
                mbar_init(empty_a + 8 * k, S3_CWARPS);
            }
            mbar_fence_init();
        }
        __syncthreads();
        int first = -1;
#pragma unroll
        for (int w = S3_WARPS - 1; w >= 0; w--)
            if (s_ball[w])
                first = w * 32 + __ffs(s_ball[w]) - 1;
        if (first < 0)
            break;
        if (tid == first)
            s_c0 = c_i;
        __syncthreads();
        const uint32_t c0 = s_c0;
        const uint32_t a_i = c_i - c0;
        const bool mine = pending && (a_i <= (uint32_t)S3_SPAN);
        const bool work = mine && active; // the lanes that take part in the pair work
        const uint32_t mine_w = __ballot_sync(0xffffffffu, mine);
        const uint32_t work_w = __ballot_sync(0xffffffffu, work);
        if (mine_w) {
            const uint32_t sp = __reduce_max_sync(0xffffffffu, mine ? a_i : 0u);
            if (lane == 0)
                atomicMax(&s_span, sp);
        }
        if (work_w && lane == 0) {
            atomicMin(&s_firstw, (uint32_t)(warp * 32 + __ffs(work_w) - 1));
            atomicMax(&s_lastw, (uint32_t)(warp * 32 + 31 - __clz(work_w)));
        }
        pending = pending && !mine;
        __syncthreads();
        if (s_firstw == 0xFFFFFFFFu)
            continue; // no particle of the group belongs to the kernel's i set
        if (tid == (int)s_firstw) {
            s_o[0] = st.x; s_o[1] = st.y; s_o[2] = st.z;
        }
        if (tid == (int)s_lastw) {
            s_o[3] = st.x; s_o[4] = st.y; s_o[5] = st.z;
        }
        const uint32_t len = s_span + 3u, nparts = (len + 1u) / 2u;
        const uint32_t NE = NROWS * nparts;
        if ((uint32_t)tid < NE) {
            const uint32_t row = tid / nparts, part = tid - row * nparts;
            const int cy = (int)(row % 3u) - 1, cz = (P::DIMS == 3) ? (int)(row / 3u) - 1 : 0;
            const uint32_t base = c0 + (uint32_t)cy * ll.nx + (uint32_t)cz * ll.nx * ll.ny - 1u;
            uint32_t lo = base + 2u * part;
            uint32_t wid = (2u * part + 1u < len) ? 1u : 0u;
            bool has0 = true, has1 = wid != 0;
            if (ll.cls) { // cells without any particle of the kernel's j classes are left out
                has0 = (ll.cls[lo] & ll.jmask) != 0;
                has1 = has1 && (ll.cls[lo + 1u] & ll.jmask) != 0;
            }
            uint32_t b = ll.N;
            if (has0)
                b = __ldg(ll.ihoc + lo);
            if (has1)
                b = min(b, __ldg(ll.ihoc + lo + 1u));
            if (!has0) {
                lo += 1u;
                wid = 0u;
            } else if (!has1) {
                wid = 0u;
            }
            uint32_t en = b;
            if (b < ll.N)
                en = s3_run_end(ll.icell, b, ll.N, lo, wid);
            else
                b = en = ll.N;
            e_begin[tid] = b;
            e_end[tid] = en;
            e_lo[tid] = lo;
            e_rel[tid] = lo - base; // x offset of cell lo relative to c0, plus one
            if (en > b)
                atomicMax(&s_maxk, (en - b + 31u) / 32u);
        }
        __syncthreads();
        const uint32_t maxk = s_maxk;
        const uint32_t nrounds = (maxk * NE + W - 1) / W;
        if (tid == 0) {
            // origin of the relative coordinates: between the first and the last working
            // member (kept in shared memory: only the staging needs it)
            s_o[0] = 0.5f * (s_o[0] + s_o[3]);
            s_o[1] = 0.5f * (s_o[1] + s_o[4]);
            s_o[2] = (P::DIMS == 3) ? 0.5f * (s_o[2] + s_o[5]) : 0.f;
            if constexpr (MODE == 1 || MODE == 3) { // rounds of this pass in the mask / count array
                uint32_t base = S3_NOPASS;
                if (npass < (uint32_t)S3_MAXPASS) {
                    const unsigned long long b = atomicAdd(pc.ctl, (unsigned long long)nrounds);
                    if (b + nrounds <= (unsigned long long)pc.cap_rounds)
                        base = (uint32_t)b;
                    pc.pass_tab[(size_t)blockIdx.x * S3_MAXPASS + npass] = base;
                } else {
                    atomicOr(pc.ctl + 1, 1ull);
                }
                s_base = base;
            } else if constexpr (MODE == 2) {
                const uint32_t base = (npass < (uint32_t)S3_MAXPASS)
                                          ? pc.pass_tab[(size_t)blockIdx.x * S3_MAXPASS + npass]
                                          : S3_NOPASS;
                if (base == S3_NOPASS && nrounds) // the host never hands over an incomplete cache
                    __trap();
                s_base = base;
            }
        }
        __syncthreads();
        // filter constants of this lane: -2 r_i and |r_i|^2 - cut^2, relative to the origin
        const float fx = -2.f * (st.x - s_o[0]), fy = -2.f * (st.y - s_o[1]);
        const float fz = (P::DIMS == 3) ? -2.f * (st.z - s_o[2]) : 0.f;
        const float fc = 0.25f * fmaf(fz, fz, fmaf(fy, fy, fx * fx)) - cut2f;

        // Per-lane FIFO of (hit mask, ring slot) entries, one per tile in which the lane has
        // hits, appended by the filter: at most NM are alive (the tiles of K - 1 rounds), so
        // qr == qw means empty.  cur = the unconsumed hits of the tile the lane is working on
        // (candidate k at bit 31 - k), cslot its ring slot, crow the shared-window address of
        // its last row slot.
        uint32_t qr = 0, qw = 0, cur = 0, cslot = 0, crow = 0;

        auto pick = [&]() { // cur == 0 && qr != qw: take the next tile with hits
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(Mw_a + qr * 128) : "memory");
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(cslot) : "r"(Sw_a + qr * 32) : "memory");
            qr = (qr + 1 == NM) ? 0u : qr + 1;
            crow = sJ_a + cslot * (SLOT4 * 16);
        };
        auto push = [&](uint32_t m, uint32_t slot) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(Mw_a + qw * 128), "r"(m) : "memory");
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(Sw_a + qw * 32), "r"(slot) : "memory");
            qw = (qw + 1 == NM) ? 0u : qw + 1;
        };
        auto body1 = [&]() { // cur != 0: the next hit of this lane
            uint32_t f;
            asm("bfind.u32 %0, %1;" : "=r"(f) : "r"(cur));
            cur ^= 1u << f;
            const uint32_t a = crow - (f << 4);
            float4 v[P::NJ4];
#pragma unroll
            for (int q = 0; q < P::NJ4; q++)
                v[q] = lds128(a + q * 512);
            if (p.test(st, v[0]))
                p.body(st, v, 1);
            if (!cur && qr != qw)
                pick();
        };
        // Two hits per iteration (P::PAIR2): both rows are loaded first and both bodies run
        // as straight-line code, so the two dependency chains (LDS -> sqrt -> Wendland
        // factors -> accumulators) overlap.  A hit that fails the exact test, or the
        // missing second hit of a lane with a single one left, gets its weights zeroed
        // (P::kill) instead of a branch: its terms are exactly +-0.
        auto body2 = [&]() { // cur != 0
            uint32_t f;
            asm("bfind.u32 %0, %1;" : "=r"(f) : "r"(cur));
            cur ^= 1u << f;
            const uint32_t a1 = crow - (f << 4);
            if (!cur && qr != qw)
                pick();
            const bool two = cur != 0;
            uint32_t a2 = a1;
            if (two) {
                asm("bfind.u32 %0, %1;" : "=r"(f) : "r"(cur));
                cur ^= 1u << f;
                a2 = crow - (f << 4);
            }
            float4 v1[P::NJ4], v2[P::NJ4];
#pragma unroll
            for (int q = 0; q < P::NJ4; q++) {
                v1[q] = lds128(a1 + q * 512);
                v2[q] = lds128(a2 + q * 512);
            }
            if (!p.test(st, v1[0]))
                P::kill(v1);
            if (!two || !p.test(st, v2[0]))
                P::kill(v2);
            p.body(st, v1, 1);
            p.body(st, v2, 1);
            if (!cur && qr != qw)
                pick();
        };
        auto consume = [&]() {
            if (cur) {
                if constexpr (P::PAIR2)
                    body2();
                else
                    body1();
            }
        };
        // stage (producer warp): S3_BATCH consecutive tiles of a round at a time; tile number tn =
        // (part e = tn % NE, its k-th tile, k = tn / NE), carried in (tk, te).  The global loads
        // of a batch are issued before any of them is used (branch-free: a lane without a
        // candidate loads particle 0 and drops it).  MODE 1 stages the test layout only, MODE 2
        // the j rows only.
        // (the builders are bound by their staging warp, which waits for one tile's loads at a time:
        // they keep S3_BBATCH tiles in flight; with pair bodies to run the consumers are the bottleneck)
        constexpr int NB = (MODE == 1 || MODE == 3) ? S3_BBATCH : S3_BATCH;
        uint32_t tk = 0, te = 0;
        auto stage_batch = [&](uint32_t ring_round, uint32_t w0) {
            const uint32_t par = ring_round;
            uint32_t cnt[NB], cj[NB], eb[NB];
            float4 o[NB][P::NJ4];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t e = te, k = tk;
                if (++te == NE) {
                    te = 0;
                    tk++;
                }
                eb[b] = e;
                const uint32_t bg = e_begin[e] + 32u * k, en = e_end[e];
                cnt[b] = (k < maxk && bg < en) ? min(32u, en - bg) : 0u;
                const uint32_t jj = ((uint32_t)lane < cnt[b]) ? bg + lane : 0u;
                if constexpr (MODE != 2)
                    cj[b] = __ldg(ll.icell + jj) - e_lo[e];
                p.stage_j(jj, o[b]);
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t w2 = w0 + b;
                uint32_t c = cnt[b];
                if (c) {
                    const bool in = (uint32_t)lane < c;
                    if constexpr (MODE != 1) {
                        float4* const slot = sJ + (size_t)(ring_round * W + w2) * SLOT4;
                        if (in) {
#pragma unroll
                            for (int q = 0; q < P::NJ4; q++)
                                slot[q * 32 + lane] = o[b][q];
                        }
                    }
                    if constexpr (MODE != 2) {
                        float tx = 0.f, ty = 0.f, tz = 0.f, tn2 = AQC_NEVER;
                        const bool live = in && P::j_live(o[b][0]);
                        if (live) {
                            tx = o[b][0].x - s_o[0];
                            ty = o[b][0].y - s_o[1];
                            tz = (P::DIMS == 3) ? o[b][0].z - s_o[2] : 0.f;
                            tn2 = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
                        }
                        float* const tst = reinterpret_cast<float*>(sT + (par * W + w2) * 32) +
                                           (lane >> 1) * 8 + (lane & 1);
                        tst[0] = tx;
                        tst[2] = ty;
                        tst[4] = tz;
                        tst[6] = tn2;
                        // a tile without a candidate that can interact at all is dropped (a tile
                        // of fluid particles costs a boundary kernel its staging only)
                        if (!__any_sync(0xffffffffu, live))
                            c = 0;
                        const uint32_t cjv = in ? cj[b] : 0xFFFFFFFFu;
                        const uint32_t cj0 = __shfl_sync(0xffffffffu, cjv, 0);
                        const int n1 = __popc(__ballot_sync(0xffffffffu, cjv == cj0));
                        if (lane == 0) {
                            t_rel[par][w2] = e_rel[eb[b]] + cj0;
                            t_n1[par][w2] = (uint32_t)n1;
                        }
                    }
                }
                if constexpr (MODE != 2) {
                    if (lane == 0)
                        t_cnt[par][w2] = c;
                }
            }
        };

        // Producer / consumers over a ring of K rounds of W tiles.  The producer warp stages
        // round r into ring round r % K as soon as every consumer warp has released the round
        // that lived there (r - K); a consumer warp filters round r when it is full, runs its
        // balanced bodies, and releases round r + 2 - K after consuming the hits it still had
        // in it.  A consumer never waits for another consumer: only the ring couples them, so
        // warps may drift apart by K - 1 rounds before anybody stalls.
        if (producer) {
            uint32_t rk = 0, use = 0; // ring round r % K, r / K
            S3P_START
            for (uint32_t r = 0; r < nrounds; r++) {
                S3P_ACC(20)
                if (use) { // ring round rk holds round r - K
                    uint32_t spins = 0;
                    while (!__all_sync(0xffffffffu, mbar_wait(empty_a + 8 * rk, (use - 1) & 1u))) {
                        if (++spins > (1u << 28)) // watchdog: a lost arrival must not hang the GPU
                            __trap();
                        if (S3_SLEEP_P)
                            __nanosleep(S3_SLEEP_P);
                    }
                }
                S3P_ACC(16)
                if constexpr (MODE == 2) {
                    // lane w2 moves tile w2 of the round: NJ4 bulk copies from the packed rows
                    // straight into the ring slot; the barrier completes with their bytes
                    uint32_t bytes = 0;
                    if (lane < W) {
                        const uint32_t tn = r * W + (uint32_t)lane;
                        const uint32_t k = tn / NE, e = tn - k * NE;
                        if (k < maxk) {
                            const uint32_t bg = e_begin[e] + 32u * k, en = e_end[e];
                            if (bg < en) {
                                bytes = min(32u, en - bg) * 16u;
                                const uint32_t dst = sJ_a - 31 * 16 + (rk * W + (uint32_t)lane) * (SLOT4 * 16);
#pragma unroll
                                for (int q = 0; q < P::NJ4; q++)
                                    bulk_g2s(dst + q * 512, pc.rows + (size_t)q * ll.N + bg, bytes, full_a + 8 * rk);
                                bytes *= P::NJ4;
                            }
                        }
                    }
                    bytes = __reduce_add_sync(0xffffffffu, bytes);
                    if (lane == 0) {
                        if (bytes)
                            mbar_arrive_expect_tx(full_a + 8 * rk, bytes);
                        else
                            mbar_arrive(full_a + 8 * rk);
                    }
                } else {
#pragma unroll 1
                    for (uint32_t w0 = 0; w0 < (uint32_t)W; w0 += NB)
                        stage_batch(rk, w0);
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(full_a + 8 * rk);
                }
                S3P_ACC(17)
                if (++rk == (uint32_t)K) {
                    rk = 0;
                    use++;
                }
            }
        } else {
            uint32_t rk = 0, use = 0;  // ring round of round r, r / K
            uint32_t rq = 0;           // ring round of round r + 2 - K, the next one to release
            const unsigned long long X2 = pack2(fx, fx), Y2 = pack2(fy, fy), Z2 = pack2(fz, fz),
                                     C2 = pack2(fc, fc);
            // MODE 3: the same test with the radius deflated by 1e-4 instead of inflated
            const unsigned long long D2s = pack2(p.cut2 * 2.0e-4f, p.cut2 * 2.0e-4f);
            (void)D2s;
            const uint32_t mbase = s_base;
#if S3_PROFILE
            const int s3p_c = work_w ? 0 : 8; // warps without a working lane: second bank
#endif
            // MODE 2: the masks of round r + 1 are requested as soon as those of round r have
            // gone to the FIFO, so their latency is covered by the pair bodies of round r
            uint32_t mk[W];
            const uint32_t* msrc = pc.masks + ((size_t)mbase * S3_TILES * S3_CWARPS + cw) * 32 + lane;
            if constexpr (MODE == 2) {
#pragma unroll
                for (int w2 = 0; w2 < W; w2++)
                    mk[w2] = (work && nrounds) ? S3_MASK_LD(msrc + w2 * (S3_CWARPS * 32)) : 0u;
            }
            S3P_START
            for (uint32_t r = 0; r < nrounds; r++) {
                S3P_ACC(s3p_c + 5)
                {
                    uint32_t spins = 0;
                    while (!__all_sync(0xffffffffu, mbar_wait(full_a + 8 * rk, use & 1u))) {
                        if (++spins > (1u << 28)) // watchdog: a lost arrival must not hang the GPU
                            __trap();
                        if (S3_SLEEP_C)
                            __nanosleep(S3_SLEEP_C);
                    }
                }
                S3P_ACC(s3p_c + 0)
                if constexpr (MODE == 1) {
                    // ---- filter only: the masks go to the cache (zero for the lanes that do
                    // not work and for the tiles that are not theirs)
                    uint32_t* dst = pc.masks + ((size_t)(mbase + r) * W * S3_CWARPS + cw) * 32 + lane;
#pragma unroll 1
                    for (int w2 = 0; w2 < W; w2++) {
                        uint32_t m = 0;
                        const uint32_t cnt = t_cnt[rk][w2];
                        if (work && cnt) {
                            const uint32_t rel = t_rel[rk][w2] - a_i;
                            const uint32_t pm = 0xFFFFFFFFu << (32u - t_n1[rk][w2]);
                            const uint32_t okm = (rel <= 2u ? pm : 0u) | (rel + 1u <= 2u ? ~pm : 0u);
                            if (okm) {
                                const float4* T = sT + (rk * W + w2) * 32;
                                if (cnt > 16)
                                    m = test_tile<16>(T, X2, Y2, Z2, C2);
                                else if (cnt > 8)
                                    m = test_tile<8>(T, X2, Y2, Z2, C2);
                                else
                                    m = test_tile<4>(T, X2, Y2, Z2, C2);
                                m &= okm;
                            }
                        }
                        __syncwarp();
                        if (mbase != S3_NOPASS)
                            S3_MASK_ST(dst + w2 * (S3_CWARPS * 32), m);
                    }
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(empty_a + 8 * rk);
                } else if constexpr (MODE == 3) {
                    // ---- filter, exact test of every selected candidate, and the survivors go
                    // to the lane's neighbour list in visiting order (tiles ascending, candidates
                    // ascending): 4 entries per chunk, the last chunk of the round padded
                    uint32_t acc0 = 0, acc1 = 0, nn = 0, nch = 0;
                    uint2* const lbase = pc.chunks + ((size_t)(blockIdx.x * S3_CWARPS + cw) * pc.capc) * 32 + lane;
                    const uint32_t ebase = (r % (uint32_t)S4_K) * (W * 512u);
                    auto flush = [&]() {
                        if (lc < pc.capc)
                            __stcs(lbase + (size_t)lc * 32, make_uint2(acc0, acc1));
                        lc++;
                        nch++;
                        acc0 = acc1 = nn = 0;
                    };
                    auto append = [&](uint32_t e) {
                        const uint32_t sh = (nn & 1u) * 16u;
                        if (nn & 2u)
                            acc1 |= e << sh;
                        else
                            acc0 |= e << sh;
                        if (++nn == 4u)
                            flush();
                    };
#pragma unroll 1
                    for (int w2 = 0; w2 < W; w2++) {
                        uint32_t m = 0, sure = 0;
                        const uint32_t cnt = t_cnt[rk][w2];
                        if (work && cnt) {
                            const uint32_t rel = t_rel[rk][w2] - a_i;
                            const uint32_t pm = 0xFFFFFFFFu << (32u - t_n1[rk][w2]);
                            const uint32_t okm = (rel <= 2u ? pm : 0u) | (rel + 1u <= 2u ? ~pm : 0u);
                            if (okm) {
                                const float4* T = sT + (rk * W + w2) * 32;
                                // the packed filter over-selects by 1e-4 of the radius; with the radius
                                // DEFLATED by as much it only under-selects (its own error is ~1e-6):
                                // those candidates need no exact test
                                if (cnt > 16)
                                    test_tile2<16>(T, X2, Y2, Z2, C2, D2s, m, sure);
                                else if (cnt > 8)
                                    test_tile2<8>(T, X2, Y2, Z2, C2, D2s, m, sure);
                                else
                                    test_tile2<4>(T, X2, Y2, Z2, C2, D2s, m, sure);
                                m &= okm;
                                sure &= m;
                            }
                        }
                        const float4* const rowp = sJ + (size_t)(rk * W + w2) * SLOT4;
                        const uint32_t etile = ebase + (uint32_t)w2 * 512u + 31u * 16u;
                        // the shell between the two radii: the readers' P::test, to the letter
                        for (uint32_t mm = m & ~sure; mm; mm &= mm - 1u) {
                            const uint32_t f = 31u - (uint32_t)__clz(mm);
                            const float4 A = rowp[31u - f];
                            if (!(dist2<P::DIMS>(A.x - st.x, A.y - st.y, A.z - st.z) < p.cut2))
                                m ^= 1u << f;
                        }
                        auto next = [&]() { // the next candidate k = 31 - f of the tile
                            const uint32_t f = 31u - (uint32_t)__clz(m);
                            m ^= 1u << f;
                            return etile - (f << 4);
                        };
                        while (m) {
                            if (nn == 0u && __popc(m) >= 4) { // a whole chunk at once
                                acc0 = next();
                                acc0 |= next() << 16;
                                acc1 = next();
                                acc1 |= next() << 16;
                                flush();
                            } else {
                                append(next());
                            }
                        }
                    }
                    if (nn) {
                        if (nn == 1u)
                            acc0 |= S4_NULL << 16;
                        if (nn <= 2u)
                            acc1 = S4_NULL | (S4_NULL << 16);
                        else
                            acc1 |= S4_NULL << 16;
                        flush();
                    }
                    if (mbase != S3_NOPASS)
                        pc.cnt[((size_t)(mbase + r) * S3_CWARPS + cw) * 32 + lane] = (uint8_t)nch;
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(empty_a + 8 * rk);
                } else {
                    if (work) {
                        if constexpr (MODE == 2) {
#pragma unroll
                            for (int w2 = 0; w2 < W; w2++)
                                if (mk[w2])
                                    push(mk[w2], rk * W + w2);
                            if (r + 1 < nrounds) {
                                msrc += W * S3_CWARPS * 32;
#pragma unroll
                                for (int w2 = 0; w2 < W; w2++)
                                    mk[w2] = S3_MASK_LD(msrc + w2 * (S3_CWARPS * 32));
                            }
                        } else {
                            // ---- filter: record the hit masks of the round's tiles
#pragma unroll 1
                            for (int w2 = 0; w2 < W; w2++) {
                                const uint32_t cnt = t_cnt[rk][w2];
                                if (!cnt)
                                    continue;
                                const uint32_t rel = t_rel[rk][w2] - a_i; // (x offset of the lower cell - a_i) + 1
                                const uint32_t pm = 0xFFFFFFFFu << (32u - t_n1[rk][w2]);
                                const uint32_t okm = (rel <= 2u ? pm : 0u) | (rel + 1u <= 2u ? ~pm : 0u);
                                if (!okm)
                                    continue;
                                const float4* T = sT + (rk * W + w2) * 32;
                                uint32_t m;
                                if (cnt > 16)
                                    m = test_tile<16>(T, X2, Y2, Z2, C2);
                                else if (cnt > 8)
                                    m = test_tile<8>(T, X2, Y2, Z2, C2);
                                else
                                    m = test_tile<4>(T, X2, Y2, Z2, C2);
                                m &= okm;
                                if (m)
                                    push(m, rk * W + w2);
                            }
                        }
                        if (!cur && qr != qw)
                            pick();
                        S3P_ACC(s3p_c + 1)
                        // ---- bodies, while every working lane of the warp has one pending
                        while (__ballot_sync(work_w, cur != 0) == work_w) {
                            if constexpr (P::PAIR2)
                                body2();
                            else
                                body1();
                        }
                    }
                    __syncwarp();
                    S3P_ACC(s3p_c + 2)
                    // ---- release round r + 2 - K: its hits are the oldest of every FIFO, a lane
                    // is done with them when the tile it works on is a newer one
                    if (r + 2 >= (uint32_t)K) {
                        while (__any_sync(0xffffffffu, cur != 0 && cslot / W == rq))
                            consume();
                        __syncwarp();
                        if (lane == 0)
                            mbar_arrive(empty_a + 8 * rq);
                        rq = (rq + 1 == (uint32_t)K) ? 0u : rq + 1;
                    }
                }
                S3P_ACC(s3p_c + 3)
                if (++rk == (uint32_t)K) {
                    rk = 0;
                    use++;
                }
            }
        }
#if S3_PROFILE
        long long s3p_t = clock64();
        const int s3p_c2 = producer ? 24 : (work_w ? 0 : 8);
#endif
        while (__any_sync(0xffffffffu, cur != 0))
            consume();
        S3P_ACC(s3p_c2 + 4)
        __syncthreads();
        S3P_ACC(s3p_c2 + 6)
#if S3_PROFILE
        if (tid == 0) {
            atomicAdd(&g_s3prof[29], (unsigned long long)nrounds);
            atomicAdd(&g_s3prof[31], 1ull);
        }
#endif
    }
    if constexpr (MODE == 3) {
        // the longest list of the warp: beyond capc the host grows the lists and builds again
        const uint32_t lmax = __reduce_max_sync(0xffffffffu, lc);
        if (lane == 0 && !producer && lmax)
            atomicMax(pc.ctl + 2, (unsigned long long)lmax);
    }
    if (active)
        p.store_i(st, i);
}

// ---------------------------------------------------------------------------
// v4 engine: the readers of the neighbour LISTS (pair cache, MODE 3 builder above).
//
// What the mask-reading sweeps (MODE 2) spent per pair besides the pair body, by their SASS
// (profiles/r1_ncu_fused_fluid_v4_maskcache_lines.txt: 157 instructions per two hits, 93 of them
// floating point): the exact re-test of every cached hit (7 + the kill selects), the decoding
// of the hit masks (find-leading-one, shift, xor, address: 4), the per-lane FIFO of (mask, slot)
// entries in shared memory (push, pick, wrap-around: ~10 issued whether they are predicated off
// or not), and the branches around the second hit of an iteration.  None of that depends on the
// kernel, so the builder does it once per geometry: it applies the exact test and stores, per
// lane, the rows to visit as a list of 16-bit ring offsets in visiting order, in chunks of four
// (the last chunk of a ring round padded with the offset of a zero-weight row), plus the number
// of chunks per (lane, round).  A reader then runs, per chunk, 2 integer instructions to split
// the two words, 4 x NJ4 LDS.128 at [ring + offset] (the ring is static shared memory: the base
// is an immediate) and four pair bodies back to back -- no test, no branch, four independent
// dependency chains.  The lane-balancing of v3 stays, at chunk granularity: a warp takes a
// chunk only while every working lane has one from the rounds in the ring, and a lane finishes
// the oldest round's chunks before the warp releases it.
// Same pairs in the same order as MODE 0 / 2 (the padding rows add exactly +-0).
#ifndef S4_MINB
#define S4_MINB 4 // CTAs per SM the register allocation aims at (kernels with two row arrays)
#endif
#ifndef S4_MINB1
#define S4_MINB1 4 // ... (kernels with one: fewer live registers, more warps pay)
#endif
#ifndef S4_SPLIT
#define S4_SPLIT 1 // 1: the rows of a chunk are fetched two at a time (fewer live registers)
#endif
#ifndef S4_NBUF
#define S4_NBUF 1 // chunks held in registers ahead of the current one (1 or 3)
#endif
#ifndef S4_SLEEP0
#define S4_SLEEP0 64 // ns: first sleep of a consumer warp that waits for its round, doubled up to
#endif
#ifndef S4_SLEEP1
#define S4_SLEEP1 1024
#endif
#ifndef S4_AHEAD
#define S4_AHEAD 6 // chunks ahead of the current one whose line is prefetched into L1 (0: none)
#endif
template <class P>
__global__ void __launch_bounds__(S3_THREADS, (P::NJ4 == 1) ? S4_MINB1 : S4_MINB)
sweep4_kernel(const P p, const LLParams ll, const S3Cache pc)
{
    constexpr int W = S3_TILES, K = S4_K;
    __shared__ float4 ring[P::NJ4][S4_ROWS + 1];
    __shared__ uint32_t e_begin[S3_MAXE], e_end[S3_MAXE];
    __shared__ uint32_t s_ball[S3_WARPS], s_c0, s_span, s_anywork, s_maxk, s_base;
    __shared__ unsigned long long s_full[K], s_empty[K];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == 0;
    const int cw = warp - 1;
    const uint32_t full_a = (uint32_t)__cvta_generic_to_shared(s_full);
    const uint32_t empty_a = (uint32_t)__cvta_generic_to_shared(s_empty);
    const uint32_t ring_a = (uint32_t)__cvta_generic_to_shared(&ring[0][0]); // (the producer's copies)
    constexpr uint32_t QSTRIDE = (S4_ROWS + 1) * 16; // bytes between the row arrays
    const uint32_t i = blockIdx.x * (uint32_t)S3_PARTICLES + (uint32_t)(tid - 32);
    const bool valid = !producer && i < ll.N;
    const bool active = valid && p.i_active(p.imove[valid ? i : 0]);
    const uint32_t c_i = valid ? __ldg(ll.icell_i + i) : 0xFFFFFFFFu;
    if constexpr (P::REMOTE) {
        if (ll.near && !__syncthreads_or(valid && c_i < ll.nw && ll.near[c_i < ll.nw ? c_i : 0]))
            return; // (the builder left this CTA at the same place)
    }
    typename P::IState st;
    st.x = st.y = st.z = 0.f;
    if (active)
        p.load_i(st, i);
    if (tid == 32) {
        // the padding row: a finite position next to the CTA's particles, every weight zero
        const float4 r0 = pc.rows[min(blockIdx.x * (uint32_t)S3_PARTICLES, ll.N - 1)];
        ring[0][S4_ROWS] = make_float4(P::j_live(r0) ? r0.x : 0.f, r0.y, r0.z, 0.f);
#pragma unroll
        for (int q = 1; q < P::NJ4; q++)
            ring[q][S4_ROWS] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int NROWS = (P::DIMS == 3) ? 9 : 3;
    // this lane's list: chunk c at lp[c * 32]
    const uint2* lp = pc.chunks + ((size_t)(blockIdx.x * S3_CWARPS + (producer ? 0 : cw)) * pc.capc) * 32 + lane;

    bool pending = valid && c_i < ll.nw;
    for (uint32_t npass = 0;; npass++) {
        // ---- the group of this pass, exactly as the builder formed it (sweep3_kernel)
        const uint32_t pb = __ballot_sync(0xffffffffu, pending);
        if (lane == 0)
            s_ball[warp] = pb;
        if (tid == 0) {
            s_span = 0;
            s_anywork = 0;
            s_maxk = 0;
            for (int k = 0; k < K; k++) {
                if (npass) {
                    mbar_inval(full_a + 8 * k);
                    mbar_inval(empty_a + 8 * k);
                }
                mbar_init(full_a + 8 * k, 1);
                mbar_init(empty_a + 8 * k, S3_CWARPS);
            }
            mbar_fence_init();
        }
        __syncthreads();
        int first = -1;
#pragma unroll
        for (int w = S3_WARPS - 1; w >= 0; w--)
            if (s_ball[w])
                first = w * 32 + __ffs(s_ball[w]) - 1;
        if (first < 0)
            break;
        if (tid == first)
            s_c0 = c_i;
        __syncthreads();
        const uint32_t c0 = s_c0;
        const uint32_t a_i = c_i - c0;
        const bool mine = pending && (a_i <= (uint32_t)S3_SPAN);
        const bool work = mine && active;
        const uint32_t mine_w = __ballot_sync(0xffffffffu, mine);
        const uint32_t work_w = __ballot_sync(0xffffffffu, work);
        if (mine_w) {
            const uint32_t sp = __reduce_max_sync(0xffffffffu, mine ? a_i : 0u);
            if (lane == 0)
                atomicMax(&s_span, sp);
        }
        if (work_w && lane == 0)
            s_anywork = 1u;
        pending = pending && !mine;
        __syncthreads();
        if (!s_anywork)
            continue; // (the builder's i classes include this kernel's: it may have worked here)
        const uint32_t len = s_span + 3u, nparts = (len + 1u) / 2u;
        const uint32_t NE = NROWS * nparts;
        if ((uint32_t)tid < NE) {
            const uint32_t row = tid / nparts, part = tid - row * nparts;
            const int cy = (int)(row % 3u) - 1, cz = (P::DIMS == 3) ? (int)(row / 3u) - 1 : 0;
            const uint32_t lo = c0 + (uint32_t)cy * ll.nx + (uint32_t)cz * ll.nx * ll.ny - 1u + 2u * part;
            const uint32_t wid = (2u * part + 1u < len) ? 1u : 0u;
            uint32_t b = __ldg(ll.ihoc + lo);
            if (wid)
                b = min(b, __ldg(ll.ihoc + lo + 1u));
            uint32_t en = b;
            if (b < ll.N)
                en = s3_run_end(ll.icell, b, ll.N, lo, wid);
            else
                b = en = ll.N;
            e_begin[tid] = b;
            e_end[tid] = en;
            if (en > b)
                atomicMax(&s_maxk, (en - b + 31u) / 32u);
        }
        __syncthreads();
        const uint32_t maxk = s_maxk;
        const uint32_t nrounds = (maxk * NE + W - 1) / W;
        if (tid == 0) {
            const uint32_t base = (npass < (uint32_t)S3_MAXPASS)
                                      ? pc.pass_tab[(size_t)blockIdx.x * S3_MAXPASS + npass]
                                      : S3_NOPASS;
            if (base == S3_NOPASS && nrounds) // the host never hands over incomplete lists
                __trap();
            s_base = base;
        }
        __syncthreads();

        if (producer) {
            // one elected lane per tile: NJ4 bulk copies from the packed rows into the ring slot,
            // completing on the round's `full` barrier
            uint32_t rk = 0, use = 0;
            for (uint32_t r = 0; r < nrounds; r++) {
                if (use) {
                    uint32_t spins = 0;
                    while (!__all_sync(0xffffffffu, mbar_wait(empty_a + 8 * rk, (use - 1) & 1u))) {
                        if (++spins > (1u << 28)) // watchdog: a lost arrival must not hang the GPU
                            __trap();
                        __nanosleep(S3_SLEEP_P);
                    }
                }
                uint32_t bytes = 0;
                if (lane < W) {
                    const uint32_t tn = r * W + (uint32_t)lane;
                    const uint32_t k = tn / NE, e = tn - k * NE;
                    if (k < maxk) {
                        const uint32_t bg = e_begin[e] + 32u * k, en = e_end[e];
                        if (bg < en) {
                            bytes = min(32u, en - bg) * 16u;
                            const uint32_t dst = ring_a + (rk * W + (uint32_t)lane) * 512u;
#pragma unroll
                            for (int q = 0; q < P::NJ4; q++)
                                bulk_g2s(dst + q * QSTRIDE, pc.rows + (size_t)q * ll.N + bg, bytes, full_a + 8 * rk);
                            bytes *= P::NJ4;
                        }
                    }
                }
                bytes = __reduce_add_sync(0xffffffffu, bytes);
                if (lane == 0) {
                    if (bytes)
                        mbar_arrive_expect_tx(full_a + 8 * rk, bytes);
                    else
                        mbar_arrive(full_a + 8 * rk);
                }
                if (++rk == (uint32_t)K) {
                    rk = 0;
                    use++;
                }
            }
        } else {
            const uint8_t* cp = pc.cnt + ((size_t)s_base * S3_CWARPS + cw) * 32 + lane;
            // chunks of this lane: consumed so far / staged in the ring so far (through the round that
            // was last found full) / belonging to the rounds up to the one to be released next
            uint32_t done = 0, avail = 0, relq = 0;
            uint32_t cn = (work && nrounds) ? (uint32_t)__ldg(cp) : 0u; // of the round to come
            uint32_t cq = cn;                                            // of the round to release next
            // the lane's list: the next chunk is loaded while the current one is worked on, and the
            // lines S4_AHEAD chunks further on are already on their way to L1 (reads past the end of
            // a list stay inside the allocation and are never used)
#if S4_NBUF == 3
            uint2 c0v = make_uint2(0u, 0u), c1v = c0v, c2v = c0v;
            if (work) {
                c0v = __ldg(lp);
                c1v = __ldg(lp + 32);
                c2v = __ldg(lp + 64);
                lp += 96;
            }
#else
            uint2 c0v = make_uint2(0u, 0u);
            if (work) {
                c0v = __ldg(lp);
                lp += 32;
            }
#endif
            // policies whose body returns early for some i particles (the fused pass: Shepard serves
            // boundary elements too) offer body_all(), the same terms without the branch: a warp
            // with at least one lane that needs everything runs it for all its lanes (what the
            // others accumulate is never stored)
            bool whole = false;
            if constexpr (P::HAS_BODY_ALL)
                whole = __any_sync(0xffffffffu, work && p.needs_all(st));
            const char* const ring_c = reinterpret_cast<const char*>(&ring[0][0]);
            auto bodies = [&](const float4* va, const float4* vb) {
                if constexpr (P::HAS_BODY_ALL) {
                    if (whole) {
                        p.body_all(st, va);
                        p.body_all(st, vb);
                        return;
                    }
                }
                p.body(st, va, 1);
                p.body(st, vb, 1);
            };
            auto take = [&]() { // one chunk: four rows, four pair bodies
                const uint2 e = c0v;
#if S4_NBUF == 3
                c0v = c1v;
                c1v = c2v;
                c2v = __ldg(lp);
#else
                c0v = __ldg(lp);
#endif
#if S4_AHEAD
                asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + S4_AHEAD * 32));
#endif
                lp += 32;
                done++;
                const uint32_t o0 = e.x & 0xFFFFu, o1 = e.x >> 16, o2 = e.y & 0xFFFFu, o3 = e.y >> 16;
                float4 v0[P::NJ4], v1[P::NJ4], v2[P::NJ4], v3[P::NJ4];
#pragma unroll
                for (int q = 0; q < P::NJ4; q++) {
                    v0[q] = *reinterpret_cast<const float4*>(ring_c + q * QSTRIDE + o0);
                    v1[q] = *reinterpret_cast<const float4*>(ring_c + q * QSTRIDE + o1);
                }
#if S4_SPLIT
                bodies(v0, v1);
                asm volatile("" ::: "memory"); // (the second half's rows are fetched after the first half's bodies)
#endif
#pragma unroll
                for (int q = 0; q < P::NJ4; q++) {
                    v2[q] = *reinterpret_cast<const float4*>(ring_c + q * QSTRIDE + o2);
                    v3[q] = *reinterpret_cast<const float4*>(ring_c + q * QSTRIDE + o3);
                }
#if !S4_SPLIT
                bodies(v0, v1);
#endif
                bodies(v2, v3);
            };
            // Round r is consumed from ring round r % K.  After finding it full a warp takes chunks
            // while every working lane has one staged; then it releases round r + 2 - K, whose chunks
            // are the oldest of every list: a lane still owing some runs them first.  Only the ring
            // couples the warps: they may drift apart by K - 2 rounds before anybody waits.
            uint32_t rk = 0, use = 0, rq = 0;
            for (uint32_t r = 0; r < nrounds; r++) {
                {
                    // a warp that finds its round not staged yet is ahead of the slowest warp of the
                    // CTA (or of the copies): it backs off, leaving the issue slots to the others
                    uint32_t spins = 0, ns = S4_SLEEP0;
                    while (!__all_sync(0xffffffffu, mbar_wait(full_a + 8 * rk, use & 1u))) {
                        if (++spins > (1u << 24)) // watchdog: a lost arrival must not hang the GPU
                            __trap();
                        __nanosleep(ns);
                        ns = min(2u * ns, (uint32_t)S4_SLEEP1);
                    }
                }
                avail += cn;
                if (work && r + 1 < nrounds)
                    cn = (uint32_t)__ldg(cp + (size_t)(r + 1) * (S3_CWARPS * 32));
                if (work) {
                    while (__ballot_sync(work_w, done < avail) == work_w)
                        take();
                }
                __syncwarp();
                if (r + 2 >= (uint32_t)K) {
                    const uint32_t q = r + 2 - (uint32_t)K;
                    relq += cq;
                    if (work)
                        cq = (uint32_t)__ldg(cp + (size_t)(q + 1) * (S3_CWARPS * 32));
                    while (__any_sync(0xffffffffu, done < relq)) {
                        if (done < relq)
                            take();
                    }
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(empty_a + 8 * rq);
                    rq = (rq + 1 == (uint32_t)K) ? 0u : rq + 1;
                }
                if (++rk == (uint32_t)K) {
                    rk = 0;
                    use++;
                }
            }
            while (__any_sync(0xffffffffu, done < avail)) {
                if (done < avail)
                    take();
            }
        }
        __syncthreads();
    }
    if (active)
        p.store_i(st, i);
}

int aqc_sweep_engine();      // 2 or 3 (AQC_SWEEP_ENGINE, default 3)
bool aqc_sweep_engine_forced(); // chosen explicitly (environment or aqc_sweep_engine_select)
int aqc_sweep_ring(int nj4); // ring rounds K of the v3 engine (AQC_SWEEP_RING)
int aqc_sweep_ring2();       // ring rounds of the mask-reading sweeps (AQC_SWEEP_RING2)
int aqc_remote_engine();     // 2 or 3 (AQC_REMOTE_ENGINE)
bool aqc_remote_lists();     // the remote sweeps read neighbour lists of their own (AQC_REMOTE_LISTS)
int aqc_pc_prepare(aqc_ctx* ctx, const int* imove, const void* r, const void* rj, int dims, float cut2,
                   const LLParams& ll, uint32_t icls, uint32_t jcls, int K, S3Cache* out); // sweeps.cu

template <class P>
static int launch_sweep(aqc_ctx* ctx, const P& p, const LLParams& ll_in)
{
    LLParams ll = ll_in;
    if (ll.N == 0)
        return AQC_OK; // nothing to do (a zero-sized grid is a launch error)
    if constexpr (P::JCLS != 0) {
        // the j set is a small class (boundary elements): rebuild the per-cell class mask
        // (imove may have changed since the last launch) and skip the cells without one
        const size_t need = ((size_t)ll.nw + 3) & ~(size_t)3;
        if (need > ctx->cell_cls_cap) {
            if (ctx->cell_cls)
                cudaFree(ctx->cell_cls);
            ctx->cell_cls = nullptr;
            ctx->cell_cls_cap = 0;
            AQC_CUDA(ctx, cudaMalloc(&ctx->cell_cls, need + need / 4));
            ctx->cell_cls_cap = need + need / 4;
        }
        AQC_CUDA(ctx, cudaMemsetAsync(ctx->cell_cls, 0, need, ctx->stream));
        cell_class_kernel<<<aqc_blocks(ll.N, 256), 256, 0, ctx->stream>>>(p.imove, ll.icell, ll.N, ll.nw,
                                                                         ctx->cell_cls);
        AQC_LAUNCH_CHECK(ctx);
        ll.cls = ctx->cell_cls;
        ll.jmask = P::JCLS;
    }
    if constexpr (P::REMOTE) {
        // remote (halo) list: flag the cells that have a halo particle in their 3^D neighbourhood
        // (AQC_REMOTE_NEAR=0: every warp walks its 27 cells, as before)
        static const bool use_near = !(getenv("AQC_REMOTE_NEAR") && atoi(getenv("AQC_REMOTE_NEAR")) == 0);
        const bool rlists = P::SPHERE && aqc_remote_lists() && ctx->pcr.enabled && aqc_sweep_engine() == 3 &&
                            !((p.icls() | p.jcls()) & ~31u);
        if ((use_near && aqc_remote_engine() != 3) || rlists) {
            const size_t need = ((size_t)ll.nw + 3) & ~(size_t)3;
            if (need > ctx->cell_cls_cap) {
                if (ctx->cell_cls) {
                    AQC_SYNC(ctx);
                    cudaFree(ctx->cell_cls);
                }
                ctx->cell_cls = nullptr;
                ctx->cell_cls_cap = 0;
                AQC_CUDA(ctx, cudaMalloc(&ctx->cell_cls, need + need / 4));
                ctx->cell_cls_cap = need + need / 4;
            }
            AQC_CUDA(ctx, cudaMemsetAsync(ctx->cell_cls, 0, need, ctx->stream));
            // (the halo rows head the sorted remote list; what follows them is the pile of
            // unused rows parked at r_max by cfd/MPI.cl::backup_r, flagged through its first row)
            remote_near_kernel<<<aqc_blocks(ll.N, 256), 256, 0, ctx->stream>>>(ll.icell, ll.N, ll.nx, ll.ny, ll.nw,
                                                                              P::DIMS, ctx->cell_cls);
            AQC_LAUNCH_CHECK(ctx);
            ll.near = ctx->cell_cls;
        }
    }
    if constexpr (P::REMOTE && P::SPHERE) {
        // remote sweeps on neighbour lists of their own (the second pair cache, ctx->pcr): built by
        // the first remote sweep after the halo list (or the local geometry) changed, read by every
        // one that follows -- inside a midpoint loop the halo positions are fixed, so the seven
        // remote sweeps of a delta-SPH step share one build
        if (ll.near && aqc_remote_lists() && ctx->pcr.enabled && aqc_sweep_engine() == 3 &&
            !((p.icls() | p.jcls()) & ~31u)) {
            S3Cache pc;
            const int cached = aqc_pc_prepare(ctx, p.imove, p.r, p.mpi_r, P::DIMS, p.cut2, ll, p.icls(), 0u,
                                              aqc_sweep_ring(1), &pc);
            if (cached < 0)
                return cached;
            if (cached == 2) {
                const size_t need = (size_t)ll.N * P::NJ4 * sizeof(float4);
                if (need > ctx->pack_cap) {
                    if (ctx->pack_rows) {
                        AQC_SYNC(ctx);
                        AQC_CUDA(ctx, cudaFree(ctx->pack_rows));
                    }
                    ctx->pack_rows = nullptr;
                    ctx->pack_cap = 0;
                    AQC_CUDA(ctx, cudaMalloc(&ctx->pack_rows, need + need / 8));
                    ctx->pack_cap = need + need / 8;
                }
                s3_pack_kernel<P><<<aqc_blocks(ll.N, 256), 256, 0, ctx->stream>>>(p, ll.N, (float4*)ctx->pack_rows);
                AQC_LAUNCH_CHECK(ctx);
                pc.rows = (const float4*)ctx->pack_rows;
                sweep4_kernel<P><<<aqc_blocks(ll.N, S3_PARTICLES), S3_THREADS, 0, ctx->stream>>>(p, ll, pc);
                AQC_LAUNCH_CHECK(ctx);
                return AQC_OK;
            }
        }
    }
    if constexpr (P::SPHERE) {
        // 2-D sweeps have ~20x less work per particle: below ~1 M particles the CTA-wide
        // set-up of v3 does not pay (measured: 2-D dam break, 0.38 M particles, 1.44 vs 1.20 ms
        // per step; 7.4 M: 12.8 vs 15.3)
        const bool small2d = (P::DIMS == 2) && ll.N < (1u << 20) && !aqc_sweep_engine_forced();
        const bool sparse = P::SPARSE_I && !(P::REMOTE && aqc_remote_engine() == 3);
        if (aqc_sweep_engine() == 3 && !sparse && !small2d) {
            int K = aqc_sweep_ring(P::NJ4);
            S3Cache pc;
            int cached = 0;
            if constexpr (P::CACHE) {
                cached = aqc_pc_prepare(ctx, p.imove, p.r, nullptr, P::DIMS, p.cut2, ll, p.icls(), p.jcls(), K, &pc);
                if (cached < 0)
                    return cached;
            }
            // shared memory: test layouts (not when the masks are read), j rows, per-lane FIFOs
            auto smem_of = [](int k, int w, bool test) {
                const size_t NS = (size_t)k * w;
                return ((test ? NS * 32 : 0) + NS * P::NJ4 * 32) * sizeof(float4) +
                       S3_CWARPS * (NS - w) * 32 * (sizeof(uint32_t) + 1);
            };
            if constexpr (P::CACHE) {
                if (cached) {
                    // the j rows of every particle, packed for the bulk copies of the readers
                    const size_t need = (size_t)ll.N * P::NJ4 * sizeof(float4);
                    if (need > ctx->pack_cap) {
                        if (ctx->pack_rows) {
                            AQC_SYNC(ctx);
                            AQC_CUDA(ctx, cudaFree(ctx->pack_rows));
                        }
                        ctx->pack_rows = nullptr;
                        ctx->pack_cap = 0;
                        AQC_CUDA(ctx, cudaMalloc(&ctx->pack_rows, need + need / 8));
                        ctx->pack_cap = need + need / 8;
                    }
                    s3_pack_kernel<P><<<aqc_blocks(ll.N, 256), 256, 0, ctx->stream>>>(p, ll.N, (float4*)ctx->pack_rows);
                    AQC_LAUNCH_CHECK(ctx);
                    pc.rows = (const float4*)ctx->pack_rows;
                }
                if (cached == 2) {
                    sweep4_kernel<P><<<aqc_blocks(ll.N, S3_PARTICLES), S3_THREADS, 0, ctx->stream>>>(p, ll, pc);
                    AQC_LAUNCH_CHECK(ctx);
                    return AQC_OK;
                }
                if (cached) {
                    K = aqc_sweep_ring2();
                    const size_t smem = smem_of(K, S3_RTILES, false);
                    static size_t configured2[64] = { 0 }; // per instantiation AND device (the attribute is per device)
                    if (smem > configured2[ctx->device & 63]) {
                        AQC_CUDA(ctx, cudaFuncSetAttribute(sweep3_kernel<P, 2, S3_RTILES>,
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        configured2[ctx->device & 63] = smem;
                    }
                    sweep3_kernel<P, 2, S3_RTILES><<<aqc_blocks(ll.N, S3_PARTICLES), S3_THREADS, smem, ctx->stream>>>(
                        p, ll, K, pc);
                    AQC_LAUNCH_CHECK(ctx);
                    return AQC_OK;
                }
            }
            const size_t smem = smem_of(K, S3_TILES, true);
            static size_t configured[64] = { 0 }; // per instantiation AND device (the attribute is per device)
            if (smem > configured[ctx->device & 63]) {
                AQC_CUDA(ctx, cudaFuncSetAttribute(sweep3_kernel<P, 0, S3_TILES>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured[ctx->device & 63] = smem;
            }
            sweep3_kernel<P, 0, S3_TILES><<<aqc_blocks(ll.N, S3_PARTICLES), S3_THREADS, smem, ctx->stream>>>(p, ll, K, pc);
        } else {
            sweep2_kernel<P><<<aqc_blocks(ll.N, SWEEP_THREADS), SWEEP_THREADS, 0, ctx->stream>>>(p, ll);
        }
    } else {
        sweep_kernel<P, true><<<aqc_blocks(ll.N, SWEEP_THREADS), SWEEP_THREADS, 0, ctx->stream>>>(p, ll);
    }
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}
