// sweep.cuh -- the neighbour sweep: one engine for every kernel the reference
// writes with BEGIN_NEIGHS/END_NEIGHS (resources/Scripts/types/3D.h:197-219,
// 2D.h:174-193).
//
// Reference shape: one work-item per particle walks the 27 (9) cell lists on
// its own, fetching icell[j], imove[j], r[j], ... from global memory per
// candidate, and recomputing per-j factors (m_j/rho_j, kernel constants) for
// every pair.
//
// B200 shape (this file):
//   * particles are cell-ordered (the link-list sort), so one warp takes 32
//     consecutive particles; lanes are grouped by cell and each group walks its
//     neighbour cells ONCE for the whole group (warp-uniform loop);
//   * the j particles of a cell are contiguous: a tile of 32 of them is loaded
//     with coalesced 16-byte loads, reduced to the few floats a pair needs
//     (position + a per-j weight such as wcon*CONF*m_j/rho_j; excluded j get a
//     far-away position) and staged in per-warp shared memory as float4 SoA;
//   * every lane tests the 32 staged candidates with broadcast LDS.128 reads
//     and records its hits in a 32-bit mask; the pair bodies then run over the
//     lane's own hits only (COMPACT), instead of the whole warp executing the
//     body whenever any lane hits;
//   * hits are visited in ascending j inside ascending cell order (x-outer, y,
//     z-inner), i.e. the reference's summation order is preserved per particle
//     (needed by the order-dependent ElasticBounce / PST kernels).
//   * no __syncthreads: warps are independent, a warp with no active particle
//     leaves immediately (sensor / boundary-only kernels).
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
};

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
                    uint32_t j0 = __ldg(ll.ihoc + cell);
                    while (j0 < ll.N) {
                        const uint32_t jj = j0 + lane;
                        const bool in = (jj < ll.N) && (__ldg(ll.icell + jj) == cell);
                        const int cnt = __popc(__ballot_sync(0xffffffffu, in));
                        if (!cnt)
                            break;
                        if (in) {
                            float4 o[P::NJ4];
                            p.stage_j(jj, o);
#pragma unroll
                            for (int k = 0; k < P::NJ4; k++)
                                tile[k][lane] = o[k];
                        }
                        __syncwarp();
                        if (mine) {
                            if constexpr (COMPACT) {
                                uint32_t hits = 0;
                                for (int k = 0; k < cnt; k++)
                                    if (p.test(st, tile[0][k]))
                                        hits |= 1u << k;
                                while (hits) {
                                    const int k = __ffs(hits) - 1;
                                    hits &= hits - 1;
                                    p.body(st, &tile[0][k], 32);
                                }
                            } else {
                                for (int k = 0; k < cnt; k++)
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

template <class P>
static int launch_sweep(aqc_ctx* ctx, const P& p, const LLParams& ll)
{
    const unsigned grid = aqc_blocks(ll.N, SWEEP_THREADS);
    if constexpr (P::SPHERE)
        sweep2_kernel<P><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(p, ll);
    else
        sweep_kernel<P, true><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(p, ll);
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}
