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
    const uint32_t c_i = active ? __ldg(ll.icell + i) : 0xFFFFFFFFu;
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

template <class P>
static int launch_sweep(aqc_ctx* ctx, const P& p, const LLParams& ll)
{
    const unsigned grid = aqc_blocks(ll.N, SWEEP_THREADS);
    sweep_kernel<P, true><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(p, ll);
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}
