// sweeps.cu -- policies of the neighbour sweep (sweep.cuh) for every
// BEGIN_NEIGHS kernel on the hot path, and their Kernel-tool registry entries.
// Argument lists (names, order) are those of the reference kernels, so the
// host binds Variables by name exactly as Kernel.cpp:497-556 does.
//
// Arithmetic notes (all inside the stated fp32 tolerance):
//  * per-j factors are hoisted to the staging step: w_j = wcon*CON*m_j/rho_j;
//  * per-i factors (1/rho_i, rho_i, __CLEARY__) are applied once after the loop;
//  * q = sqrt(d2)/H is evaluated as sqrt.approx(d2)*(1/H) (1 ulp); the filter is
//    d2 < (SUPPORT*H)^2 instead of q >= SUPPORT (differs only for pairs within
//    an ulp of the cut-off, where the Wendland factors (2-q)^3,(2-q)^4 vanish);
//  * the i == j exclusion of the reference is implicit wherever the pair term is
//    exactly zero for r_ij = 0 (all kernels below that exclude it).
#include "sweep.cuh"

#include <stdlib.h>

// Engine selection for the SPHERE policies (A/B measurements; both engines are
// CUDA): AQC_SWEEP_ENGINE=2 keeps the immediate-body engine, default 3 = deferred.
static int g_sweep_engine = 0;
static bool g_sweep_forced = false;
int aqc_sweep_engine()
{
    if (!g_sweep_engine) {
        const char* s = getenv("AQC_SWEEP_ENGINE");
        g_sweep_forced = s && (atoi(s) == 2 || atoi(s) == 3);
        g_sweep_engine = (s && atoi(s) == 2) ? 2 : 3;
    }
    return g_sweep_engine;
}
bool aqc_sweep_engine_forced()
{
    aqc_sweep_engine();
    return g_sweep_forced;
}
extern "C" int aqc_sweep_engine_select(int engine)
{
    if (engine == 2 || engine == 3) {
        g_sweep_engine = engine;
        g_sweep_forced = true;
    } else if (engine < 0) { // back to the automatic choice (environment, then size heuristics)
        g_sweep_engine = 0;
        g_sweep_forced = false;
    }
    return aqc_sweep_engine();
}
#if S3_PROFILE
// debug builds only: per-phase clock totals of sweep3_kernel (see S3P_ACC in sweep.cuh)
extern "C" int aqc_debug_s3prof(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    if (out)
        cudaMemcpyFromSymbol(out, g_s3prof, sizeof(g_s3prof));
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(g_s3prof, z, sizeof(z));
    }
    return 0;
}
#endif
// Ring rounds of the v3 engine (8 tiles each): shared memory per CTA =
// 8 KB + K * 8 * NJ4 * 512 + (K - 1) * 8 * 1024 bytes; the deferral window is (K - 1) * 8 tiles
int aqc_sweep_ring(int nj4)
{
    static const int forced = [] {
        const char* s = getenv("AQC_SWEEP_RING");
        const int r = s ? atoi(s) : 0;
        return (r >= 2 && r <= 8) ? r : 0;
    }();
    if (forced)
        return forced;
    (void)nj4;
    return 3;
}

int aqc_remote_engine() // engine of the remote (halo) sweeps: 2 (per warp, default) or 3 (CTA rings)
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AQC_REMOTE_ENGINE");
        v = (e && atoi(e) == 3) ? 3 : 2;
    }
    return v;
}

bool aqc_remote_lists() // the remote (halo) sweeps read neighbour lists of their own (AQC_REMOTE_LISTS, default on)
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AQC_REMOTE_LISTS");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
}

int aqc_sweep_ring2()
{
    static const int forced = [] {
        const char* s = getenv("AQC_SWEEP_RING2");
        const int r = s ? atoi(s) : 0;
        return (r >= 2 && r <= S3_MAXK) ? r : 0;
    }();
    return forced ? forced : S3_RRING;
}

namespace {

constexpr float iM_PI = 0.318309886f; // KernelFunctions/Wendland3D.hcl:34-39

template <int DIMS> struct Wend; // Wendland{2D,3D}.hcl:44-66
template <> struct Wend<3> {
    static constexpr float W = 0.08203125f * iM_PI;
    static constexpr float F = 0.8203125f * iM_PI;
    static constexpr float CLEARY = 10.f; // cfd/Interactions.cl:33-39
};
template <> struct Wend<2> {
    static constexpr float W = 0.109375f * iM_PI;
    static constexpr float F = 1.09375f * iM_PI;
    static constexpr float CLEARY = 8.f;
};

template <int DIMS> struct VecT;
template <> struct VecT<3> { using T = float4; };
template <> struct VecT<2> { using T = float2; };

template <int DIMS>
__device__ __forceinline__ float4 ldvec(const void* base, uint32_t i)
{
    if constexpr (DIMS == 3) {
        return __ldg(reinterpret_cast<const float4*>(base) + i);
    } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(base) + i);
        return make_float4(t.x, t.y, 0.f, 0.f);
    }
}
// plain (coherent) load for arrays the same kernel also writes
template <int DIMS>
__device__ __forceinline__ float4 ldvec_rw(const void* base, uint32_t i)
{
    if constexpr (DIMS == 3) {
        return reinterpret_cast<const float4*>(base)[i];
    } else {
        const float2 t = reinterpret_cast<const float2*>(base)[i];
        return make_float4(t.x, t.y, 0.f, 0.f);
    }
}
// write only the XYZ (XY) components, like "v[i].XYZ = ..." in the reference
template <int DIMS>
__device__ __forceinline__ void stvec_xyz(void* base, uint32_t i, float x, float y, float z)
{
    if constexpr (DIMS == 3) {
        float* p = reinterpret_cast<float*>(base) + 4 * (size_t)i;
        p[0] = x; p[1] = y; p[2] = z;
    } else {
        reinterpret_cast<float2*>(base)[i] = make_float2(x, y);
    }
}

// q from d2 (see header note).  sqrt.approx.ftz is one MUFU.SQRT with a maximum
// relative error of 2^-23 (1 ulp) and maps 0 -> 0, so no special case is needed.
__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 1 / x, one MUFU.RCP (maximum relative error 2^-23)
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float q_of(float d2, float invH)
{
    return sqrt_fast(d2) * invH;
}

struct PBase {
    const int* __restrict__ imove;
    float invH, cut2;
    // staged row 0 of a j that can interact at all (tiles without one are skipped)
    __device__ static bool j_live(const float4& o0) { return o0.x != AQC_FAR; }
    // kernels whose i particles are few and scattered (sensors, boundary elements) keep
    // the per-warp engine: a CTA-wide walk for a handful of lanes does not pay
    static constexpr bool SPARSE_I = false;
    // classes (aqc_cls_bit) the j particles belong to when they are a small minority (boundary
    // elements): the launcher then builds a per-cell class mask and cells without one are skipped
    static constexpr uint32_t JCLS = 0;
    // v3 engine: two hits per body iteration.  Needs kill() to zero every weight of a
    // staged row so that body() adds exactly +-0 for it.
    static constexpr bool PAIR2 = false;
    __device__ static void kill(float4* v) { v[0].w = 0.f; }
    // v3 engine: the kernel may read the pair-mask cache (sweep.cuh, S3Cache).  icls() / jcls()
    // = classes (aqc_cls_bit) its i and j particles can belong to; they must cover i_active()
    // and the candidates stage_j() leaves alive.
    static constexpr bool CACHE = false;
    // v4 engine: body_all() / needs_all() exist (see PFusedFluid)
    static constexpr bool HAS_BODY_ALL = false;
    // the j list is a remote (halo) one: SPARSE_I by default, the CTA engine with AQC_REMOTE_ENGINE=3
    static constexpr bool REMOTE = false;
    uint32_t icls() const { return 0xFFu; }
    uint32_t jcls() const { return 0xFFu; }
};

// ------------------------------------------------------------------------
// cfd/Interactions.cl:60-145
template <int D>
struct PInteractions : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = true;
    uint32_t icls() const { return 1u; }
    uint32_t jcls() const { return 1u; }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *u;
    const float *rho, *m, *p;
    void *grad_p, *lap_u;
    float* div_u;
    float cF;     // wconF * CONF
    float eps2;   // 0.01 * H * H
    struct IState { float x, y, z, ux, uy, uz, p, gx, gy, gz, lx, ly, lz, du; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.p = __ldg(p + i);
        s.gx = s.gy = s.gz = s.lx = s.ly = s.lz = s.du = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == 1;
        const float w = cF * __ldg(m + j) / __ldg(rho + j);
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, w);
        o[1] = make_float4(b.x, b.y, b.z, __ldg(p + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float t = 2.f - q_of(d2, invH);
        const float fr = (t * t) * (t * A.w); // kernelF(q)*CONF*m_j / rho_j
        float udr = (B.x - s.ux) * dx + (B.y - s.uy) * dy;
        if constexpr (D == 3)
            udr += (B.z - s.uz) * dz;
        const float a = (s.p + B.w) * fr;
        const float b0 = udr * fr;
        const float b = b0 * rcp_fast(d2 + eps2);
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
        s.du += b0;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float cl = Wend<D>::CLEARY * ir;
        stvec_xyz<D>(grad_p, i, s.gx * ir, s.gy * ir, s.gz * ir);
        stvec_xyz<D>(lap_u, i, s.lx * cl, s.ly * cl, s.lz * cl);
        div_u[i] = s.du * rho_i;
    }
};

// cfd/Interactions.cl:60-145 compiled with __LAP_FORMULATION__ = __LAP_MORRIS__ (:130-131; the <Define> of
// examples/2D/taylor_green and cylinder_inside_channel): lap_u takes f_ij * 2 / (rho_i rho_j) * (u_j - u_i)
// instead of the Cleary term along r_ij; grad_p and div_u are the Monaghan build's.  Never fused
// (aqc_launch_fused refuses under this definition): the fused groups are built for the default.
template <int D>
struct PInteractionsMorris : PInteractions<D> {
    using IState = typename PInteractions<D>::IState;
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float t = 2.f - q_of(d2, this->invH);
        const float fr = (t * t) * (t * A.w); // kernelF(q)*CONF*m_j / rho_j
        const float dux = B.x - s.ux, duy = B.y - s.uy, duz = B.z - s.uz;
        float udr = dux * dx + duy * dy;
        if constexpr (D == 3)
            udr += duz * dz;
        const float a = (s.p + B.w) * fr;
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        s.lx += fr * dux; s.ly += fr * duy;
        if constexpr (D == 3)
            s.lz += fr * duz;
        s.du += udr * fr;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float rho_i = __ldg(this->rho + i);
        const float ir = 1.f / rho_i;
        const float c2 = 2.f * ir;
        stvec_xyz<D>(this->grad_p, i, s.gx * ir, s.gy * ir, s.gz * ir);
        stvec_xyz<D>(this->lap_u, i, s.lx * c2, s.ly * c2, s.lz * c2);
        this->div_u[i] = s.du * rho_i;
    }
};

// ------------------------------------------------------------------------
// basic/Shepard.cl:76-125 (MODE 0) and cfd/Shepard.cl:29-35 (MODE 1)
template <int D, int MODE>
struct PShepard : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = MODE == 1; // (MODE 0 also serves imove == 2)
    uint32_t icls() const { return 31u; }
    uint32_t jcls() const { return 1u; }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void* r;
    const float *rho, *m;
    float* shepard;
    float cW; // wconW * CONW
    struct IState { float x, y, z, s; };
    __device__ static bool excl(int mv) { return MODE ? (mv != 1) : (mv >= 3); }
    __device__ bool i_active(int mv) const { return !((mv < -3) || ((mv > 0) && excl(mv))); }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.s = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        const bool ok = !excl(__ldg(imove + j));
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j) / __ldg(rho + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        s.s += (1.f + 2.f * q) * (t2 * t2) * A.w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const { shepard[i] = s.s; }
};

// ------------------------------------------------------------------------
// cfd/Boundary/Portal/Shepard.cl:44-113 and Portal/Interactions.cl:47-146 (preset cfd/portal.xml): the particles
// Portal/Mirror.cl::mirror has moved to the in portal (imirrored) ADD what they see there to their sums.  A
// mirrored particle keeps its row of the sorted arrays while icell[i] names the cell it fell into, so these
// sweeps stay on the per-warp engine, whose lanes take their cell from icell[i] (SPARSE_I; the mirrored
// particles are few as well).  PBase::imove points at imirrored -- the engine's i filter -- and the
// particle class is read from `mv`.
template <int D>
struct PPortalShepard : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool SPARSE_I = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const int* mv; // imove
    const void* r;
    const float *rho, *m;
    float* shepard;
    float cW; // wconW * CONW
    struct IState { float x, y, z, s; bool ok; };
    __device__ bool i_active(int mirrored) const { return mirrored != 0; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        const int c = __ldg(mv + i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.s = 0.f;
        s.ok = !((c < -3) || (c > 1));
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        const bool ok = (__ldg(mv + j) == 1) && !__ldg(imove + j);
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j) / __ldg(rho + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        s.s += (1.f + 2.f * q) * (t2 * t2) * A.w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if (s.ok)
            shepard[i] = shepard[i] + s.s;
    }
};

template <int D, bool MORRIS>
struct PPortalInteractions : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool SPARSE_I = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const int* mv; // imove
    const void *r, *u;
    const float *rho, *m, *p;
    void *grad_p, *lap_u;
    float* div_u;
    float cF;     // wconF * CONF
    float eps2;   // 0.01 * H * H
    struct IState { float x, y, z, ux, uy, uz, p, gx, gy, gz, lx, ly, lz, du; bool ok; };
    __device__ bool i_active(int mirrored) const { return mirrored != 0; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.p = __ldg(p + i);
        s.gx = s.gy = s.gz = s.lx = s.ly = s.lz = s.du = 0.f;
        s.ok = __ldg(mv + i) == 1;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec<D>(u, j);
        const int c = __ldg(mv + j);
        const bool ok = !__ldg(imove + j) && (c == 1 || c == -1);
        const float w = cF * __ldg(m + j) / __ldg(rho + j);
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, w);
        o[1] = make_float4(b.x, b.y, b.z, __ldg(p + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float t = 2.f - q_of(d2, invH);
        const float fr = (t * t) * (t * A.w); // kernelF(q)*CONF*m_j / rho_j
        const float dux = B.x - s.ux, duy = B.y - s.uy, duz = B.z - s.uz;
        float udr = dux * dx + duy * dy;
        if constexpr (D == 3)
            udr += duz * dz;
        const float a = (s.p + B.w) * fr;
        const float b0 = udr * fr;
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        if constexpr (MORRIS) {
            s.lx += fr * dux; s.ly += fr * duy; s.lz += fr * duz;
        } else {
            const float b = b0 * rcp_fast(d2 + eps2);
            s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
        }
        s.du += b0;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if (!s.ok)
            return;
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float cl = (MORRIS ? 2.f : Wend<D>::CLEARY) * ir;
        const float4 g = ldvec<D>(grad_p, i), l = ldvec<D>(lap_u, i);
        stvec_xyz<D>(grad_p, i, g.x + s.gx * ir, g.y + s.gy * ir, g.z + s.gz * ir);
        stvec_xyz<D>(lap_u, i, l.x + s.lx * cl, l.y + s.ly * cl, l.z + s.lz * cl);
        div_u[i] = div_u[i] + s.du * rho_i;
    }
};

// ------------------------------------------------------------------------
// basic/deltaSPH.cl:94-145 (full, VECOUT) and :191-242 (lapp); EXCLUDED = imove != 1
template <int D, bool VECOUT>
struct PDeltaGrad : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = true;
    uint32_t icls() const { return 1u; }
    uint32_t jcls() const { return 1u; }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void* r;
    const float *rho, *m, *p;
    void* out; // lap_p_corr (vec) or lap_p (float)
    float cF;
    struct IState { float x, y, z, p, ax, ay, az; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.p = __ldg(p + i);
        s.ax = s.ay = s.az = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        const bool ok = __ldg(imove + j) == 1;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cF * __ldg(m + j) / __ldg(rho + j));
        o[1] = make_float4(__ldg(p + j), 0.f, 0.f, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0];
        const float pj = row[stride].x;
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float t = 2.f - q_of(dist2<D>(dx, dy, dz), invH);
        const float c = (pj - s.p) * ((t * t) * (t * A.w));
        if constexpr (VECOUT) {
            s.ax += c * dx; s.ay += c * dy; s.az += c * dz;
        } else {
            s.ax += c;
        }
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if constexpr (VECOUT)
            stvec_xyz<D>(out, i, s.ax, s.ay, s.az);
        else
            reinterpret_cast<float*>(out)[i] = s.ax;
    }
};

// basic/deltaSPH.cl:261-313 (lapp_corr): starts from the old lap_p[i] (:287)
template <int D>
struct PLappCorr : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = true;
    uint32_t icls() const { return 1u; }
    uint32_t jcls() const { return 1u; }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void* r;
    const float *rho, *m;
    const void* lap_p_corr;
    float* lap_p;
    float cF;
    struct IState { float x, y, z, gx, gy, gz, acc; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), g = ldvec<D>(lap_p_corr, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.gx = g.x; s.gy = g.y; s.gz = g.z;
        s.acc = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), g = ldvec<D>(lap_p_corr, j);
        const bool ok = __ldg(imove + j) == 1;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cF * __ldg(m + j) / __ldg(rho + j));
        o[1] = make_float4(g.x, g.y, g.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], G = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float t = 2.f - q_of(dist2<D>(dx, dy, dz), invH);
        float gr = (G.x + s.gx) * dx + (G.y + s.gy) * dy;
        if constexpr (D == 3)
            gr += (G.z + s.gz) * dz;
        s.acc += gr * ((t * t) * (t * A.w));
    }
    __device__ void store_i(const IState& s, uint32_t i) const { lap_p[i] -= 0.5f * s.acc; }
};

// ------------------------------------------------------------------------
// basic/MLS.cl:58-112
template <int D>
struct PMLS : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = true;
    uint32_t icls() const { return aqc_cls_bit((int)mls_imove); }
    uint32_t jcls() const { return aqc_cls_bit((int)mls_imove); }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void* r;
    const float *rho, *m;
    float* mls;
    uint32_t mls_imove;
    float cF;
    struct IState { float x, y, z, a[D * D]; };
    __device__ bool i_active(int mv) const { return (uint32_t)mv == mls_imove; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
#pragma unroll
        for (int k = 0; k < D * D; k++)
            s.a[k] = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        const bool ok = (uint32_t)__ldg(imove + j) == mls_imove;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cF * __ldg(m + j) / __ldg(rho + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        float d[3] = { A.x - s.x, A.y - s.y, A.z - s.z };
        const float t = 2.f - q_of(dist2<D>(d[0], d[1], d[2]), invH);
        const float f = (t * t) * (t * A.w);
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int b = 0; b < D; b++)
                s.a[a * D + b] += d[a] * (f * d[b]);
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if constexpr (D == 3) {
            float4* o = reinterpret_cast<float4*>(mls) + 4 * (size_t)i;
            o[0] = make_float4(s.a[0], s.a[1], s.a[2], 0.f);
            o[1] = make_float4(s.a[3], s.a[4], s.a[5], 0.f);
            o[2] = make_float4(s.a[6], s.a[7], s.a[8], 0.f);
            o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            reinterpret_cast<float4*>(mls)[i] = make_float4(s.a[0], s.a[1], s.a[2], s.a[3]);
        }
    }
};

// ------------------------------------------------------------------------
// cfd/Sensors.cl:57-130
template <int D>
struct PSensors : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool SPARSE_I = true;
    static constexpr int DIMS = D, NJ4 = 3;
    const void* r;
    const float* m;
    void* u;
    float *rho, *p;
    float cW;
    struct IState { float x, y, z, ux, uy, uz, rho, p; };
    __device__ bool i_active(int mv) const { return mv == 0; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.ux = s.uy = s.uz = s.rho = s.p = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec_rw<D>(u, j);
        const bool ok = __ldg(imove + j) == 1;
        const float rj = rho[j];
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j) / rj);
        o[1] = make_float4(b.x, b.y, b.z, p[j]);
        o[2] = make_float4(rj, 0.f, 0.f, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float rj = row[2 * stride].x;
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        const float w = (1.f + 2.f * q) * (t2 * t2) * A.w;
        s.ux += B.x * w; s.uy += B.y * w; s.uz += B.z * w;
        s.rho += rj * w;
        s.p += B.w * w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        stvec_xyz<D>(u, i, s.ux, s.uy, s.uz);
        rho[i] = s.rho;
        p[i] = s.p;
    }
};

// ------------------------------------------------------------------------
// cfd/Boundary/BIe/Interactions.cl:48-108
template <int D>
struct PBIeInteractions : PBase {
    static constexpr bool SPHERE = true;
    static constexpr uint32_t JCLS = 16u;
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *normal, *u;
    const float* m;
    void* grad_w_bi;
    float* div_u_bi;
    float cW;
    struct IState { float x, y, z, gx, gy, gz, du; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.gx = s.gy = s.gz = s.du = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), n = ldvec<D>(normal, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == -3;
        float un = b.x * n.x + b.y * n.y;
        if constexpr (D == 3)
            un += b.z * n.z;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j));
        o[1] = make_float4(n.x, n.y, n.z, un);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        const float w = (1.f + 2.f * q) * (t2 * t2) * A.w; // kernelW*CONW*area_j
        s.gx += Nn.x * w; s.gy += Nn.y * w; s.gz += Nn.z * w;
        s.du -= Nn.w * w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        stvec_xyz<D>(grad_w_bi, i, s.gx, s.gy, s.gz);
        div_u_bi[i] = s.du;
    }
};

// cfd/Boundary/BIe/Interactions.cl:124-170 (p_boundary)
template <int D>
struct PBIePBoundary : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool SPARSE_I = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void* r;
    const float *m, *rho;
    float* p;
    float cW;
    struct IState { float x, y, z, p; };
    __device__ bool i_active(int mv) const { return mv == -3; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.p = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        const bool ok = __ldg(imove + j) == 1;
        // p is read for fluid j and written for boundary i: disjoint rows
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z,
                           ok ? 2.f * p[j] * cW * __ldg(m + j) / __ldg(rho + j) : 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        s.p += (1.f + 2.f * q) * (t2 * t2) * A.w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const { p[i] = s.p; }
};

// cfd/Boundary/BIe/ElasticBounce.cl:64-152 -- order dependent
template <int D>
struct PBIeElasticBounce : PBase {
    static constexpr bool SPHERE = false;
    static constexpr uint32_t JCLS = 16u;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r_in, *normal, *u_in;
    const float* m;
    void* dudt;
    float dt, dr_factor, min_bound;
    struct IState { float x, y, z, u0x, u0y, u0z, ax, ay, az, Ux, Uy, Uz; };
    __device__ bool i_active(int mv) const { return mv == 1 && dt != 0.f; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r_in, i), b = ldvec<D>(u_in, i), c = ldvec_rw<D>(dudt, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.u0x = b.x; s.u0y = b.y; s.u0z = b.z;
        s.ax = c.x; s.ay = c.y; s.az = c.z;
        s.Ux = s.u0x + 0.5f * dt * s.ax;
        s.Uy = s.u0y + 0.5f * dt * s.ay;
        s.Uz = s.u0z + 0.5f * dt * s.az;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r_in, j), n = ldvec<D>(normal, j);
        const bool ok = __ldg(imove + j) == -3;
        const float mj = __ldg(m + j);
        const float dr = (D == 3) ? sqrtf(mj) : mj;
        const float R = dr_factor * dr; // __DR_FACTOR__ (:31-33)
        o[0] = make_float4(a.x, a.y, a.z, ok ? R * R : -1.f);
        o[1] = make_float4(n.x, n.y, n.z, dr);
    }
    __device__ static bool j_live(const float4& o0) { return o0.w >= 0.f; }
    __device__ bool test(const IState& s, const float4& A) const
    {
        // cheap necessary condition: |rt|^2 < R^2 implies |r_ij|^2 - rn^2 < R^2; the
        // exact tests run in body().  Excluded j carry R^2 = -1.
        return A.w >= 0.f;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = (D == 3) ? A.z - s.z : 0.f;
        float rn = dx * Nn.x + dy * Nn.y;
        if constexpr (D == 3)
            rn += dz * Nn.z;
        if (rn < 0.f)
            return;
        const float tx = dx - rn * Nn.x, ty = dy - rn * Nn.y, tz = dz - rn * Nn.z;
        if (dist2<D>(tx, ty, tz) >= A.w)
            return;
        float Un = s.Ux * Nn.x + s.Uy * Nn.y;
        if constexpr (D == 3)
            Un += s.Uz * Nn.z;
        const float drn = dt * Un;
        if (drn < 0.f)
            return;
        if (rn - drn <= min_bound * Nn.w) { // __MIN_BOUND_DIST__ (:34-36)
            const float ux = s.u0x + dt * s.ax, uy = s.u0y + dt * s.ay, uz = s.u0z + dt * s.az;
            float un = ux * Nn.x + uy * Nn.y;
            if constexpr (D == 3)
                un += uz * Nn.z;
            const float rx = ux - 2.f * un * Nn.x, ry = uy - 2.f * un * Nn.y,
                        rz = uz - 2.f * un * Nn.z;
            s.ax = (rx - s.u0x) / dt; s.ay = (ry - s.u0y) / dt; s.az = (rz - s.u0z) / dt;
            s.Ux = s.u0x + 0.5f * dt * s.ax;
            s.Uy = s.u0y + 0.5f * dt * s.ay;
            s.Uz = s.u0z + 0.5f * dt * s.az;
        }
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        stvec_xyz<D>(dudt, i, s.ax, s.ay, s.az);
    }
};

// cfd/Boundary/BIe/PST.cl:62-110 -- order dependent (r_i moves inside the loop)
template <int D>
struct PBIePST : PBase {
    static constexpr bool SPHERE = false;
    static constexpr uint32_t JCLS = 16u;
    static constexpr int DIMS = D, NJ4 = 2;
    void* r;
    const void* normal;
    const float *m, *rho;
    float inv_dims; // 1.f / DIMS
    float dr_factor;
    struct IState { float x, y, z, Ri; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec_rw<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.Ri = 0.5f * powf(__ldg(m + i) / __ldg(rho + i), inv_dims);
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        // r is read for boundary j (never moved: only imove == 1 rows are written)
        const float4 a = ldvec_rw<D>(r, j), n = ldvec<D>(normal, j);
        const bool ok = __ldg(imove + j) == -3;
        const float mj = __ldg(m + j);
        const float dr = (D == 3) ? sqrtf(mj) : mj;
        const float R = dr_factor * dr;
        o[0] = make_float4(a.x, a.y, a.z, ok ? R * R : -1.f);
        o[1] = make_float4(n.x, n.y, n.z, 0.f);
    }
    __device__ static bool j_live(const float4& o0) { return o0.w >= 0.f; }
    __device__ bool test(const IState&, const float4& A) const { return A.w >= 0.f; }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = (D == 3) ? A.z - s.z : 0.f;
        float rn = dx * Nn.x + dy * Nn.y;
        if constexpr (D == 3)
            rn += dz * Nn.z;
        if (fabsf(rn) > s.Ri)
            return;
        const float tx = dx - rn * Nn.x, ty = dy - rn * Nn.y, tz = dz - rn * Nn.z;
        if (dist2<D>(tx, ty, tz) >= A.w)
            return;
        const float k = rn - s.Ri;
        s.x += k * Nn.x; s.y += k * Nn.y;
        if constexpr (D == 3)
            s.z += k * Nn.z;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        stvec_xyz<D>(r, i, s.x, s.y, s.z);
    }
};

// ------------------------------------------------------------------------
// cfd/MPI.cl:328-374 (gamma) and :404-485 (interactions): the halo particles
// received from the neighbour processes have their own link-list on the local
// grid; every remote j counts (no imove test) and the result is ADDED to what
// the local sweeps left in the output arrays.
template <int D>
struct PMpiGamma : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true; // the remote list is a thin halo: most CTAs find nothing
    uint32_t icls() const { return 31u; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 1;
    const void *r, *mpi_r;
    const float *mpi_rho, *mpi_m;
    float* shepard;
    float cW;
    struct IState { float x, y, z, s; };
    __device__ bool i_active(int mv) const { return !((mv < -3) || ((mv > 0) && (mv != 1))); }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.s = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j);
        o[0] = make_float4(a.x, a.y, a.z, cW * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        s.s += (1.f + 2.f * q) * (t2 * t2) * A.w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const { shepard[i] += s.s; }
};

template <int D>
struct PMpiInteractions : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true; // the remote list is a thin halo: most CTAs find nothing
    uint32_t icls() const { return 1u; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *u, *mpi_r, *mpi_u;
    const float *rho, *p, *mpi_rho, *mpi_p, *mpi_m;
    void *grad_p, *lap_u;
    float* div_u;
    float cF, eps2;
    struct IState { float x, y, z, ux, uy, uz, p, gx, gy, gz, lx, ly, lz, du; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.p = __ldg(p + i);
        s.gx = s.gy = s.gz = s.lx = s.ly = s.lz = s.du = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j), b = ldvec<D>(mpi_u, j);
        o[0] = make_float4(a.x, a.y, a.z, cF * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
        o[1] = make_float4(b.x, b.y, b.z, __ldg(mpi_p + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float t = 2.f - q_of(d2, invH);
        const float fr = (t * t) * (t * A.w);
        float udr = (B.x - s.ux) * dx + (B.y - s.uy) * dy;
        if constexpr (D == 3)
            udr += (B.z - s.uz) * dz;
        const float a = (s.p + B.w) * fr;
        const float b0 = udr * fr;
        const float b = b0 * rcp_fast(d2 + eps2);
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
        s.du += b0;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float cl = Wend<D>::CLEARY * ir;
        const float4 g0 = ldvec_rw<D>(grad_p, i), l0 = ldvec_rw<D>(lap_u, i);
        stvec_xyz<D>(grad_p, i, g0.x + s.gx * ir, g0.y + s.gy * ir, g0.z + s.gz * ir);
        stvec_xyz<D>(lap_u, i, l0.x + s.lx * cl, l0.y + s.ly * cl, l0.z + s.lz * cl);
        div_u[i] += s.du * rho_i;
    }
};

// cfd/MPI.cl interactions + gamma in one pass over the remote (halo) list: the two kernels
// share i set (gamma's is the wider one), candidates and geometry.  Same expressions as the
// members; the Shepard weight cW m_j/rho_j is derived from the staged cF m_j/rho_j.
template <int D, bool DELTA>
struct PMpiFused : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true;
    uint32_t icls() const { return 31u; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *u, *mpi_r, *mpi_u;
    const float *rho, *p, *mpi_rho, *mpi_p, *mpi_m;
    void *grad_p, *lap_u, *lap_p_corr;
    float *div_u, *shepard, *lap_p;
    float cF, cWF, eps2;
    struct IState {
        float x, y, z, ux, uy, uz, p, gx, gy, gz, lx, ly, lz, du, sh, ax, ay, az, lp;
        bool fluid;
    };
    __device__ bool i_active(int mv) const { return !((mv < -3) || ((mv > 0) && (mv != 1))); }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.fluid = __ldg(imove + i) == 1;
        s.ux = s.uy = s.uz = s.p = 0.f;
        if (s.fluid) {
            const float4 b = ldvec<D>(u, i);
            s.ux = b.x; s.uy = b.y; s.uz = b.z;
            s.p = __ldg(p + i);
        }
        s.gx = s.gy = s.gz = s.lx = s.ly = s.lz = s.du = s.sh = 0.f;
        s.ax = s.ay = s.az = s.lp = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j), b = ldvec<D>(mpi_u, j);
        o[0] = make_float4(a.x, a.y, a.z, cF * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
        o[1] = make_float4(b.x, b.y, b.z, __ldg(mpi_p + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float q = q_of(d2, invH);
        const float t = 2.f - q, t2 = t * t;
        s.sh += (1.f + 2.f * q) * (t2 * t2) * (A.w * cWF);
        if (!s.fluid)
            return;
        const float4 B = row[stride];
        const float fr = t2 * (t * A.w);
        if constexpr (DELTA) { // aqua/MPIdeltaSPH.cl::full_lapp (PMpiDelta::body)
            const float c = (B.w - s.p) * fr;
            s.ax += c * dx; s.ay += c * dy; s.az += c * dz;
            s.lp += c;
        }
        float udr = (B.x - s.ux) * dx + (B.y - s.uy) * dy;
        if constexpr (D == 3)
            udr += (B.z - s.uz) * dz;
        const float a = (s.p + B.w) * fr;
        const float b0 = udr * fr;
        const float b = b0 * rcp_fast(d2 + eps2);
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
        s.du += b0;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        shepard[i] += s.sh;
        if (!s.fluid)
            return;
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float cl = Wend<D>::CLEARY * ir;
        const float4 g0 = ldvec_rw<D>(grad_p, i), l0 = ldvec_rw<D>(lap_u, i);
        stvec_xyz<D>(grad_p, i, g0.x + s.gx * ir, g0.y + s.gy * ir, g0.z + s.gz * ir);
        stvec_xyz<D>(lap_u, i, l0.x + s.lx * cl, l0.y + s.ly * cl, l0.z + s.lz * cl);
        div_u[i] += s.du * rho_i;
        if constexpr (DELTA) {
            const float4 c0 = ldvec_rw<D>(lap_p_corr, i);
            stvec_xyz<D>(lap_p_corr, i, c0.x + s.ax, c0.y + s.ay, c0.z + s.az);
            lap_p[i] += s.lp;
        }
    }
};

// ------------------------------------------------------------------------
// Remote (halo) terms of MLS and delta-SPH: aqua/MPIdeltaSPH.cl::mls / ::full_lapp / ::lapp_corr.
// NOT reference kernels.  The reference's MPI preset exchanges what cfd/Interactions.cl and the
// Shepard factor need and nothing else (resources/Presets/src/cfd/MPI.xml:59-87), so its
// multi-process runs cannot carry delta-SPH or MLS; the slab pipelines of this repository that do
// (casegen.slab_delta_sph) insert these next to their local twins.  Each is the j loop of the local
// script (basic/MLS.cl:58-112, basic/deltaSPH.cl:94-145 + 191-242, 261-313) over the halo list,
// written the way cfd/MPI.cl:328-485 writes its own: every halo particle counts, the result is
// ADDED to what the local kernel left.
template <int D>
struct PMpiMLS : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true;
    uint32_t icls() const { return (mls_imove == 1u) ? 1u : 0xFFu; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 1;
    const void *r, *mpi_r;
    const float *mpi_rho, *mpi_m;
    float* mls;
    uint32_t mls_imove;
    float cF;
    struct IState { float x, y, z, a[D * D]; };
    __device__ bool i_active(int mv) const { return (uint32_t)mv == mls_imove; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
#pragma unroll
        for (int k = 0; k < D * D; k++)
            s.a[k] = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j);
        o[0] = make_float4(a.x, a.y, a.z, cF * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int) const
    {
        const float4 A = row[0];
        float d[3] = { A.x - s.x, A.y - s.y, A.z - s.z };
        const float t = 2.f - q_of(dist2<D>(d[0], d[1], d[2]), invH);
        const float f = (t * t) * (t * A.w);
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int b = 0; b < D; b++)
                s.a[a * D + b] += d[a] * (f * d[b]);
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if constexpr (D == 3) {
            float4* o = reinterpret_cast<float4*>(mls) + 4 * (size_t)i;
            float4 m0 = o[0], m1 = o[1], m2 = o[2];
            m0.x += s.a[0]; m0.y += s.a[1]; m0.z += s.a[2];
            m1.x += s.a[3]; m1.y += s.a[4]; m1.z += s.a[5];
            m2.x += s.a[6]; m2.y += s.a[7]; m2.z += s.a[8];
            o[0] = m0; o[1] = m1; o[2] = m2;
        } else {
            float4 m0 = reinterpret_cast<float4*>(mls)[i];
            m0.x += s.a[0]; m0.y += s.a[1]; m0.z += s.a[2]; m0.w += s.a[3];
            reinterpret_cast<float4*>(mls)[i] = m0;
        }
    }
};

template <int D>
struct PMpiDelta : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true;
    uint32_t icls() const { return 1u; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *mpi_r;
    const float *p, *mpi_rho, *mpi_m, *mpi_p;
    void* lap_p_corr;
    float* lap_p;
    float cF;
    struct IState { float x, y, z, p, ax, ay, az, lp; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.p = __ldg(p + i);
        s.ax = s.ay = s.az = s.lp = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j);
        o[0] = make_float4(a.x, a.y, a.z, cF * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
        o[1] = make_float4(__ldg(mpi_p + j), 0.f, 0.f, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0];
        const float pj = row[stride].x;
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float t = 2.f - q_of(dist2<D>(dx, dy, dz), invH);
        const float c = (pj - s.p) * ((t * t) * (t * A.w));
        s.ax += c * dx; s.ay += c * dy; s.az += c * dz;
        s.lp += c;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float4 g0 = ldvec_rw<D>(lap_p_corr, i);
        stvec_xyz<D>(lap_p_corr, i, g0.x + s.ax, g0.y + s.ay, g0.z + s.az);
        lap_p[i] += s.lp;
    }
};

template <int D>
struct PMpiLappCorr : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr bool SPARSE_I = true;
    uint32_t icls() const { return 1u; } // (the remote neighbour lists, ctx->pcr)
    uint32_t jcls() const { return 0u; }
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *mpi_r, *lap_p_corr, *mpi_lap_p_corr;
    const float *mpi_rho, *mpi_m;
    float* lap_p;
    float cF;
    struct IState { float x, y, z, gx, gy, gz, acc; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), g = ldvec<D>(lap_p_corr, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.gx = g.x; s.gy = g.y; s.gz = g.z;
        s.acc = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(mpi_r, j), g = ldvec<D>(mpi_lap_p_corr, j);
        o[0] = make_float4(a.x, a.y, a.z, cF * __ldg(mpi_m + j) / __ldg(mpi_rho + j));
        o[1] = make_float4(g.x, g.y, g.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], G = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float t = 2.f - q_of(dist2<D>(dx, dy, dz), invH);
        float gr = (G.x + s.gx) * dx + (G.y + s.gy) * dy;
        if constexpr (D == 3)
            gr += (G.z + s.gz) * dz;
        s.acc += gr * ((t * t) * (t * A.w));
    }
    __device__ void store_i(const IState& s, uint32_t i) const { lap_p[i] -= 0.5f * s.acc; }
};

// ------------------------------------------------------------------------
// Boundary integrals, cfd/Boundary/BI/*.cl (2-D dam break, BASELINE config 1)

// KernelFunctions/Wendland{2D,3D}.hcl:107-185: analytic Shepard terms of a flat element
template <int D> __device__ __forceinline__ float kernelS_P(float q);
template <> __device__ __forceinline__ float kernelS_P<2>(float q)
{
    const float wcon = 0.109375f * iM_PI;
    const float q2 = q * q, q3 = q2 * q;
    return wcon * (0.285714f * q3 * q2 - 2.5f * q2 * q2 + 8.f * q3 - 10.f * q2 + 8.f);
}
template <> __device__ __forceinline__ float kernelS_P<3>(float q)
{
    const float wcon = 0.08203125f * iM_PI;
    const float q2 = q * q, q3 = q2 * q;
    return wcon * (0.25f * q3 * q2 - 2.142857f * q2 * q2 + 6.666667f * q3 - 8.f * q2 + 5.333333f);
}
__device__ __forceinline__ float omega3(float a, float b)
{
    const float a2 = a * a, b2 = b * b;
    const float v = sqrtf((1.f + a2 + b2) / ((1.f + a2) * (1.f + b2)));
    return acosf(v < 1.f ? v : 1.f) * iM_PI; // acospi(min(.., 1))
}
__device__ __forceinline__ float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : x); }
template <int D> __device__ __forceinline__ float kernelS_D(float d, float t, float b, float s);
template <> __device__ __forceinline__ float kernelS_D<2>(float d, float t, float, float s)
{
    const float dr = 0.5f * s;
    return -(0.5f * iM_PI) * (atanf((t + dr) / d) - atanf((t - dr) / d));
}
template <> __device__ __forceinline__ float kernelS_D<3>(float d, float t, float b, float s)
{
    const float dr = 0.5f * sqrtf(s);
    const float t1 = (t - dr) / d, t2 = (t + dr) / d, b1 = (b - dr) / d, b2 = (b + dr) / d;
    const float st = signf(t1), sb = signf(b1);
    return -0.25f * (omega3(t2, b2) - st * omega3(t1, b2) - sb * omega3(t2, b1) +
                     st * sb * omega3(t1, b1));
}

// BI/Shepard.cl:61-150 (compute)
template <int D>
struct PBIShepard : PBase {
    static constexpr bool SPHERE = true;
    static constexpr uint32_t JCLS = 16u;
    static constexpr int DIMS = D, NJ4 = 4;
    const void *r, *normal, *tangent, *binormal;
    const float* m;
    float* shepard;
    float H, CONW, inv_dm1; // 1 / (DIMS - 1)
    struct IState { float x, y, z, s; uint32_t i; bool self; };
    __device__ bool i_active(int mv) const { return !((mv < -3) || (mv > 1)); }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.s = 1.f; s.i = i; s.self = false;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), n = ldvec<D>(normal, j), t = ldvec<D>(tangent, j),
                     b = ldvec<D>(binormal, j);
        const bool ok = __ldg(imove + j) == -3;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, __ldg(m + j));
        o[1] = make_float4(n.x, n.y, n.z, __uint_as_float(j));
        o[2] = make_float4(t.x, t.y, t.z, 0.f);
        o[3] = make_float4(b.x, b.y, b.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride], T = row[2 * stride], B = row[3 * stride];
        if (__float_as_uint(Nn.w) == s.i) { // a boundary element meeting itself
            if (!s.self) {
                s.self = true;
                s.s -= 0.5f;
            }
            return;
        }
        const float dx = A.x - s.x, dy = A.y - s.y, dz = (D == 3) ? A.z - s.z : 0.f;
        const float q = sqrtf(dist2<D>(dx, dy, dz)) / H;
        float rn = dx * Nn.x + dy * Nn.y, rt = dx * T.x + dy * T.y, rb = dx * B.x + dy * B.y;
        if constexpr (D == 3) {
            rn += dz * Nn.z; rt += dz * T.z; rb += dz * B.z;
        }
        rt = fabsf(rt);
        rb = fabsf(rb);
        const float area = A.w;
        if ((rn > -1e-8f * H) && (rn < 1e-8f * H)) { // lying on the boundary
            if (!s.self) {
                const float dr = 0.55f * ((D == 3) ? powf(area, inv_dm1) : area);
                if ((rt <= dr) && (rb <= dr)) {
                    s.self = true;
                    s.s -= 0.5f;
                }
            }
            return;
        }
        s.s += rn * CONW * kernelS_P<D>(q) * area + kernelS_D<D>(fabsf(rn), rt, rb, area);
    }
    __device__ void store_i(const IState& s, uint32_t i) const { shepard[i] = s.s; }
};

// BI/LapU.cl:58-123 (freeslip): boundary elements gather the fluid's velocity Laplacian
template <int D>
struct PBILapU : PBase {
    static constexpr bool SPHERE = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *u;
    const float *rho, *m;
    void* lap_u;
    float cF, eps2;
    struct IState { float x, y, z, ux, uy, uz, lx, ly, lz; };
    __device__ bool i_active(int mv) const { return mv == -3; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.lx = s.ly = s.lz = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == 1;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z,
                           cF * Wend<D>::CLEARY * __ldg(m + j) / __ldg(rho + j));
        o[1] = make_float4(b.x, b.y, b.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float t = 2.f - q_of(d2, invH);
        float udr = (B.x - s.ux) * dx + (B.y - s.uy) * dy;
        if constexpr (D == 3)
            udr += (B.z - s.uz) * dz;
        const float b = udr * ((t * t) * (t * A.w)) * rcp_fast(d2 + eps2);
        s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
    }
    __device__ void store_i(const IState& s, uint32_t i) const { stvec_xyz<D>(lap_u, i, s.lx, s.ly, s.lz); }
};

// BI/Interpolation.cl:54-107: pressure of the boundary elements from the fluid,
// corrected with the element's own pressure gradient
template <int D>
struct PBIInterpolation : PBase {
    static constexpr bool SPHERE = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *grad_p;
    const float *m, *rho;
    float* p;
    float cW;
    struct IState { float x, y, z, gx, gy, gz, p; };
    __device__ bool i_active(int mv) const { return mv == -3; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), g = ldvec<D>(grad_p, i);
        const float ri = __ldg(rho + i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.gx = ri * g.x; s.gy = ri * g.y; s.gz = ri * g.z;
        s.p = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        // p is read for fluid j and written for boundary i: disjoint rows
        const float4 a = ldvec<D>(r, j);
        const bool ok = __ldg(imove + j) == 1;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j) / __ldg(rho + j));
        o[1] = make_float4(ok ? p[j] : 0.f, 0.f, 0.f, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0];
        const float pj = row[stride].x;
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float q = q_of(dist2<D>(dx, dy, dz), invH);
        const float t = 2.f - q, t2 = t * t;
        const float w = (1.f + 2.f * q) * (t2 * t2) * A.w;
        float gr = s.gx * dx + s.gy * dy;
        if constexpr (D == 3)
            gr += s.gz * dz;
        s.p += (pj - gr) * w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const { p[i] = s.p; }
};

// BI/Interactions.cl:51-133: fluid i against boundary elements j; starts from the values the
// fluid-fluid sweep left in grad_p / div_u (:92-93)
template <int D>
struct PBIInteractions : PBase {
    static constexpr bool SPHERE = true;
    static constexpr uint32_t JCLS = 16u;
    static constexpr int DIMS = D, NJ4 = 3;
    const void *r, *normal, *u;
    const float *rho, *m, *p;
    void* grad_p;
    float* div_u;
    float cW;
    struct IState { float x, y, z, ux, uy, uz, p, gx, gy, gz, du; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.p = __ldg(p + i);
        s.gx = s.gy = s.gz = s.du = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), n = ldvec<D>(normal, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == -3;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j));
        o[1] = make_float4(n.x, n.y, n.z, __ldg(p + j));
        o[2] = make_float4(b.x, b.y, b.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride], U = row[2 * stride];
        const float q = q_of(dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z), invH);
        const float t = 2.f - q, t2 = t * t;
        const float w = (1.f + 2.f * q) * (t2 * t2) * A.w; // kernelW*CONW*area_j
        const float a = (s.p + Nn.w) * w;
        float dun = (U.x - s.ux) * Nn.x + (U.y - s.uy) * Nn.y;
        if constexpr (D == 3)
            dun += (U.z - s.uz) * Nn.z;
        s.gx += a * Nn.x; s.gy += a * Nn.y; s.gz += a * Nn.z;
        s.du += dun * w;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float4 g0 = ldvec_rw<D>(grad_p, i);
        stvec_xyz<D>(grad_p, i, g0.x + s.gx * ir, g0.y + s.gy * ir, g0.z + s.gz * ir);
        div_u[i] += rho_i * s.du;
    }
};

// BI/NoSlip.cl:52-130 (preset cfd/BINoSlip.xml): fluid i against the boundary elements of the set
// noslip_iset; adds to the lap_u the fluid-fluid sweep left (__LAP_MONAGHAN__: the Cleary term along
// n_j plus the tangential velocity difference over the wall distance, never closer than dr)
template <int D>
struct PBINoSlip : PBase {
    static constexpr bool SPHERE = true;
    static constexpr uint32_t JCLS = 16u;
    static constexpr int DIMS = D, NJ4 = 3;
    const uint32_t* iset;
    const void *r, *normal, *u;
    const float *rho, *m;
    void* lap_u;
    uint32_t noslip_iset;
    float cW, H2, dr;
    struct IState { float x, y, z, ux, uy, uz, irho, lx, ly, lz; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.irho = 1.f / __ldg(rho + i);
        s.lx = s.ly = s.lz = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), n = ldvec<D>(normal, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == -3 && __ldg(iset + j) == noslip_iset;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cW * __ldg(m + j));
        o[1] = make_float4(n.x, n.y, n.z, 0.f);
        o[2] = make_float4(b.x, b.y, b.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride], U = row[2 * stride];
        const float rx = A.x - s.x, ry = A.y - s.y, rz = D == 3 ? A.z - s.z : 0.f;
        const float q = q_of(dist2<D>(rx, ry, rz), invH);
        const float t = 2.f - q, t2 = t * t;
        const float w = (1.f + 2.f * q) * (t2 * t2) * A.w * s.irho; // kernelW*CONW*area_j / rho_i
        const float dux = U.x - s.ux, duy = U.y - s.uy, duz = D == 3 ? U.z - s.uz : 0.f;
        float dudr = dux * rx + duy * ry, rn = rx * Nn.x + ry * Nn.y, dun = dux * Nn.x + duy * Nn.y;
        if constexpr (D == 3) {
            dudr += duz * rz;
            rn += rz * Nn.z;
            dun += duz * Nn.z;
        }
        const float c1 = (D == 3 ? 10.f : 8.f) * w * dudr / ((q * q + 0.01f) * H2);
        const float c2 = 2.f * w / fmaxf(fabsf(rn), dr);
        s.lx += c1 * Nn.x + c2 * (dux - dun * Nn.x);
        s.ly += c1 * Nn.y + c2 * (duy - dun * Nn.y);
        if constexpr (D == 3)
            s.lz += c1 * Nn.z + c2 * (duz - dun * Nn.z);
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float4 l0 = ldvec_rw<D>(lap_u, i);
        stvec_xyz<D>(lap_u, i, l0.x + s.lx, l0.y + s.ly, l0.z + s.lz);
    }
};

// cfd/ideal_gas/riemann/Interactions.cl:50-168 (examples/2D/shock_1d, shock_point_riemann): the acoustic
// Riemann solver between fluid particles -- along the line of centres l_ij the pair meets at the star
// state (u*, p*), which replaces the particle averages in the continuity, momentum and energy sums.
// Per j: c_j = 2 m_j / (rho_j H) * wconF * CONW and rs_j = rho_j s_j are staged; per i the sums are
// divided / multiplied by rho_i once, at the end.  The self pair (l_ij undefined) is left out by the
// test, like the script's i == j.
template <int D>
struct PRiemann : PBase {
    static constexpr bool SPHERE = true;
    static constexpr int DIMS = D, NJ4 = 3;
    const uint32_t* iset;
    const void *r, *u;
    const float *rho, *m, *p, *gamma;
    void* grad_p;
    float *div_u, *work_density;
    float cF; // 2 / H * wconF * CONW
    struct IState { float x, y, z, ux, uy, uz, p, rs, gx, gy, gz, du, wd; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ float rs_of(uint32_t k) const
    {
        const float rho_k = __ldg(rho + k);
        return rho_k * sqrtf(__ldg(gamma + __ldg(iset + k)) * __ldg(p + k) / rho_k); // sound_speed.hcl:22-25
    }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec<D>(u, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.p = __ldg(p + i);
        s.rs = rs_of(i);
        s.gx = s.gy = s.gz = s.du = s.wd = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == 1;
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cF * __ldg(m + j) / __ldg(rho + j));
        o[1] = make_float4(b.x, b.y, b.z, __ldg(p + j));
        o[2] = make_float4(ok ? rs_of(j) : 1.f, 0.f, 0.f, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        const float d2 = dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z);
        return d2 < cut2 && d2 > 0.f;
    }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], B = row[stride];
        const float rs_j = row[2 * stride].x;
        const float dx = A.x - s.x, dy = A.y - s.y, dz = D == 3 ? A.z - s.z : 0.f;
        const float d2 = dist2<D>(dx, dy, dz);
        const float il = 1.f / sqrtf(d2);
        const float lx = dx * il, ly = dy * il, lz = dz * il;
        const float q = sqrtf(d2) * invH;
        float u_R_i = s.ux * lx + s.uy * ly, u_R_j = B.x * lx + B.y * ly;
        if constexpr (D == 3) {
            u_R_i += s.uz * lz;
            u_R_j += B.z * lz;
        }
        const float inv = 1.f / (rs_j + s.rs);
        const float u_star = (u_R_j * rs_j + u_R_i * s.rs - B.w + s.p) * inv;
        const float p_star = (B.w * s.rs + s.p * rs_j - rs_j * s.rs * (u_R_j - u_R_i)) * inv;
        const float t = 2.f - q;
        const float w = -q * ((t * t) * (t * A.w)); // c_j * W'_ij
        const float aux = (u_R_i - u_star) * w;
        const float g = p_star * w;
        s.du += aux;
        s.gx -= g * lx; s.gy -= g * ly; s.gz -= g * lz;
        s.wd += p_star * aux;
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        stvec_xyz<D>(grad_p, i, s.gx * ir, s.gy * ir, s.gz * ir);
        div_u[i] = s.du * rho_i;
        work_density[i] = s.wd * ir;
    }
};

// cfd/Boundary/ElasticBounce.cl:77-148 -- order dependent (u_i, dudt_i change inside the loop)
template <int D>
struct PElasticBounce : PBase {
    static constexpr bool SPHERE = false;
    static constexpr uint32_t JCLS = 8u | 16u;
    static constexpr int DIMS = D, NJ4 = 4;
    const void *r, *normal;
    void *u, *dudt;
    float dt, R2, min_dist, one_plus_e;
    struct IState { float x, y, z, ux, uy, uz, ax, ay, az; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i), b = ldvec_rw<D>(u, i), c = ldvec_rw<D>(dudt, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.ux = b.x; s.uy = b.y; s.uz = b.z;
        s.ax = c.x; s.ay = c.y; s.az = c.z;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        // u / dudt are read for boundary j and written for fluid i: disjoint rows
        const float4 a = ldvec<D>(r, j), n = ldvec<D>(normal, j), b = ldvec_rw<D>(u, j),
                     c = ldvec_rw<D>(dudt, j);
        const int mv = __ldg(imove + j);
        o[0] = make_float4(a.x, a.y, a.z, (mv == -2 || mv == -3) ? 1.f : -1.f);
        o[1] = make_float4(n.x, n.y, n.z, 0.f);
        o[2] = make_float4(b.x, b.y, b.z, 0.f);
        o[3] = make_float4(c.x, c.y, c.z, 0.f);
    }
    __device__ static bool j_live(const float4& o0) { return o0.w >= 0.f; }
    __device__ bool test(const IState&, const float4& A) const { return A.w >= 0.f; }
    __device__ void body(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0], Nn = row[stride], U = row[2 * stride], Acc = row[3 * stride];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = (D == 3) ? A.z - s.z : 0.f;
        float r0 = dx * Nn.x + dy * Nn.y;
        if constexpr (D == 3)
            r0 += dz * Nn.z;
        if (r0 < 0.f)
            return;
        const float tx = dx - r0 * Nn.x, ty = dy - r0 * Nn.y, tz = dz - r0 * Nn.z;
        float rt2 = tx * tx + ty * ty;
        if constexpr (D == 3)
            rt2 += tz * tz;
        if (rt2 >= R2)
            return;
        float un = (s.ux - U.x) * Nn.x + (s.uy - U.y) * Nn.y;
        float an = (s.ax - Acc.x) * Nn.x + (s.ay - Acc.y) * Nn.y;
        if constexpr (D == 3) {
            un += (s.uz - U.z) * Nn.z;
            an += (s.az - Acc.z) * Nn.z;
        }
        const float dist = dt * un + 0.5f * dt * dt * an;
        if (dist < 0.f)
            return;
        if (r0 - dist <= min_dist) {
            const float ka = one_plus_e * an, ku = one_plus_e * un;
            s.ax -= ka * Nn.x; s.ay -= ka * Nn.y;
            s.ux -= ku * Nn.x; s.uy -= ku * Nn.y;
            if constexpr (D == 3) {
                s.az -= ka * Nn.z;
                s.uz -= ku * Nn.z;
            }
        }
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        stvec_xyz<D>(u, i, s.ux, s.uy, s.uz);
        stvec_xyz<D>(dudt, i, s.ax, s.ay, s.az);
    }
};

// ------------------------------------------------------------------------
// FUSED fluid sweep.  The reference runs cfd/Shepard, cfd/Interactions, deltaSPH::full
// and deltaSPH::lapp as four separate neighbour sweeps over the SAME (i, j) pairs
// (j fluid), each re-reading r, imove, icell and recomputing |r_ij|, q and the
// kernel factors (SURVEY 2.4 "fusion targets").  Here one pass filters the
// candidates once and evaluates every member's pair term from the shared
// geometry.  Each member's arithmetic is exactly that of its stand-alone policy
// (same expressions, same pairs, same order), so the outputs equal those of running the
// members one after the other up to the compiler's FMA contraction (a few ulp).  The host asks for a fusion with
// aqc_fused_lookup(); members: Interactions always, Shepard / full / lapp optional.
template <int D, bool SHEP, bool FULL, bool LAPP>
struct PFusedFluid : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool CACHE = true;
    uint32_t icls() const { return SHEP ? 31u : 1u; }
    uint32_t jcls() const { return 1u; }
    static constexpr bool PAIR2 = true;
    static constexpr int DIMS = D, NJ4 = 2;
    const void *r, *u;
    const float *rho, *m, *p;
    void *grad_p, *lap_u, *lap_p_corr;
    float *div_u, *shepard, *lap_p;
    float cF, cW, cWF, eps2; // cWF = cW / cF
    struct IState {
        float x, y, z, ux, uy, uz, p, gx, gy, gz, lx, ly, lz, du, sh, cx, cy, cz, lp;
        bool fluid;
    };
    // union of the members' i sets: Shepard also serves sensors and boundary elements
    __device__ bool i_active(int mv) const
    {
        return SHEP ? !((mv < -3) || ((mv > 0) && (mv != 1))) : (mv == 1);
    }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
        s.fluid = !SHEP || (__ldg(imove + i) == 1);
        s.ux = s.uy = s.uz = s.p = 0.f;
        if (s.fluid) {
            const float4 b = ldvec<D>(u, i);
            s.ux = b.x; s.uy = b.y; s.uz = b.z;
            s.p = __ldg(p + i);
        }
        s.gx = s.gy = s.gz = s.lx = s.ly = s.lz = s.du = s.sh = s.cx = s.cy = s.cz = s.lp = 0.f;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j), b = ldvec<D>(u, j);
        const bool ok = __ldg(imove + j) == 1;
        const float mj = __ldg(m + j), rj = __ldg(rho + j);
        o[0] = make_float4(ok ? a.x : AQC_FAR, a.y, a.z, cF * mj / rj);
        o[1] = make_float4(b.x, b.y, b.z, __ldg(p + j));
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    // v4 engine: every term without the early return of the non-fluid i particles (their
    // u = p = 0 give finite terms that store_i never writes)
    static constexpr bool HAS_BODY_ALL = SHEP;
    __device__ bool needs_all(const IState& s) const { return s.fluid; }
    __device__ void body_all(IState& s, const float4* row) const { body_t<false>(s, row, 1); }
    __device__ void body(IState& s, const float4* row, int stride) const { body_t<true>(s, row, stride); }
    template <bool EARLY>
    __device__ __forceinline__ void body_t(IState& s, const float4* row, int stride) const
    {
        const float4 A = row[0];
        const float dx = A.x - s.x, dy = A.y - s.y, dz = A.z - s.z;
        const float d2 = dist2<D>(dx, dy, dz);
        const float q = q_of(d2, invH);
        const float t = 2.f - q;
        if constexpr (SHEP) {
            const float t2 = t * t;
            s.sh += (1.f + 2.f * q) * (t2 * t2) * (A.w * cWF); // cW m_j/rho_j from the staged cF m_j/rho_j
            if constexpr (EARLY) {
                if (!s.fluid)
                    return;
            }
        }
        const float4 B = row[stride];
        const float fr = (t * t) * (t * A.w); // kernelF(q)*CONF*m_j / rho_j
        float udr = (B.x - s.ux) * dx + (B.y - s.uy) * dy;
        if constexpr (D == 3)
            udr += (B.z - s.uz) * dz;
        const float a = (s.p + B.w) * fr;
        const float b0 = udr * fr;
        const float b = b0 * rcp_fast(d2 + eps2);
        s.gx += a * dx; s.gy += a * dy; s.gz += a * dz;
        s.lx += b * dx; s.ly += b * dy; s.lz += b * dz;
        s.du += b0;
        if constexpr (FULL || LAPP) {
            const float c = (B.w - s.p) * fr;
            if constexpr (FULL) {
                s.cx += c * dx; s.cy += c * dy; s.cz += c * dz;
            }
            if constexpr (LAPP)
                s.lp += c;
        }
    }
    __device__ void store_i(const IState& s, uint32_t i) const
    {
        if constexpr (SHEP)
            shepard[i] = s.sh;
        if (!s.fluid)
            return;
        const float rho_i = __ldg(rho + i);
        const float ir = 1.f / rho_i;
        const float cl = Wend<D>::CLEARY * ir;
        stvec_xyz<D>(grad_p, i, s.gx * ir, s.gy * ir, s.gz * ir);
        stvec_xyz<D>(lap_u, i, s.lx * cl, s.ly * cl, s.lz * cl);
        div_u[i] = s.du * rho_i;
        if constexpr (FULL)
            stvec_xyz<D>(lap_p_corr, i, s.cx, s.cy, s.cz);
        if constexpr (LAPP)
            lap_p[i] = s.lp;
    }
};

// ------------------------------------------------------------------------
// Diagnostic (not a reference script): number of fluid neighbours within the kernel
// support of every fluid particle, i.e. the pair count the roofline figures use.
template <int D>
struct PCountPairs : PBase {
    static constexpr bool SPHERE = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void* r;
    uint32_t* n_pairs;
    struct IState { float x, y, z; uint32_t n; };
    __device__ bool i_active(int mv) const { return mv == 1; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z; s.n = 0;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        o[0] = make_float4(__ldg(imove + j) == 1 ? a.x : AQC_FAR, a.y, a.z, 0.f);
    }
    __device__ bool test(const IState& s, const float4& A) const
    {
        return dist2<D>(A.x - s.x, A.y - s.y, A.z - s.z) < cut2;
    }
    __device__ void body(IState& s, const float4*, int) const { s.n++; }
    __device__ void store_i(const IState& s, uint32_t i) const { n_pairs[i] = s.n; }
};

// ------------------------------------------------------------------------
// Builder of the pair-mask cache (sweep3_kernel MODE 1): the candidate filter alone, for the i
// particles of the classes icl against the j particles of the classes jcl.
template <int D>
struct PMaskBuild : PBase {
    static constexpr bool SPHERE = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void* r;
    uint32_t icl, jcl;
    struct IState { float x, y, z; };
    __device__ bool i_active(int mv) const { return (aqc_cls_bit(mv) & icl) != 0; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(r, j);
        o[0] = make_float4((aqc_cls_bit(__ldg(imove + j)) & jcl) ? a.x : AQC_FAR, a.y, a.z, 0.f);
    }
    __device__ bool test(const IState&, const float4&) const { return false; }
    __device__ void body(IState&, const float4*, int) const {}
    __device__ void store_i(const IState&, uint32_t) const {}
};

// ... of the remote (halo) sweeps: i particles of the classes icl from the local arrays, every row of
// the halo list as j (cfd/MPI.cl:328-485 let every halo particle count)
template <int D>
struct PMaskBuildR : PBase {
    static constexpr bool SPHERE = true;
    static constexpr bool REMOTE = true;
    static constexpr int DIMS = D, NJ4 = 1;
    const void *r, *rj;
    uint32_t icl;
    struct IState { float x, y, z; };
    __device__ bool i_active(int mv) const { return (aqc_cls_bit(mv) & icl) != 0; }
    __device__ void load_i(IState& s, uint32_t i) const
    {
        const float4 a = ldvec<D>(r, i);
        s.x = a.x; s.y = a.y; s.z = a.z;
    }
    __device__ void stage_j(uint32_t j, float4* o) const
    {
        const float4 a = ldvec<D>(rj, j);
        o[0] = make_float4(a.x, a.y, a.z, 0.f);
    }
    __device__ bool test(const IState&, const float4&) const { return false; }
    __device__ void body(IState&, const float4*, int) const {}
    __device__ void store_i(const IState&, uint32_t) const {}
};

template <class PB>
int pc_build_launch(aqc_ctx* ctx, aqc_pair_cache& c, const PB& p, const LLParams& ll, int K, const S3Cache& pc)
{
    const size_t NS = (size_t)K * S3_TILES;
    const size_t smem = (NS * 32 + NS * 32) * sizeof(float4) + S3_CWARPS * (NS - S3_TILES) * 32 * (sizeof(uint32_t) + 1);
    const unsigned grid = aqc_blocks(ll.N, S3_PARTICLES);
    if (c.lists) {
        AQC_CUDA(ctx, cudaFuncSetAttribute(sweep3_kernel<PB, 3, S3_TILES>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep3_kernel<PB, 3, S3_TILES><<<grid, S3_THREADS, smem, ctx->stream>>>(p, ll, K, pc);
    } else {
        AQC_CUDA(ctx, cudaFuncSetAttribute(sweep3_kernel<PB, 1, S3_TILES>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep3_kernel<PB, 1, S3_TILES><<<grid, S3_THREADS, smem, ctx->stream>>>(p, ll, K, pc);
    }
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

template <int D>
int pc_build(aqc_ctx* ctx, aqc_pair_cache& c, const LLParams& ll, int K)
{
    S3Cache pc;
    pc.masks = c.masks;
    pc.pass_tab = c.pass_tab;
    pc.ctl = c.ctl;
    pc.cap_rounds = (uint32_t)c.cap_rounds;
    pc.chunks = (uint2*)c.chunks;
    pc.cnt = c.cnt;
    pc.capc = c.capc;
    if (c.remote) {
        PMaskBuildR<D> p;
        p.imove = (const int*)c.imove;
        p.invH = 0.f;
        p.cut2 = c.cut2;
        p.r = c.r;
        p.rj = c.rj;
        p.icl = c.icls_want;
        return pc_build_launch(ctx, c, p, ll, K, pc);
    }
    PMaskBuild<D> p;
    p.imove = (const int*)c.imove;
    p.invH = 0.f;
    p.cut2 = c.cut2;
    p.r = c.r;
    p.icl = c.icls_want;
    p.jcl = c.jcls_want;
    return pc_build_launch(ctx, c, p, ll, K, pc);
}

} // namespace

// 1: *out describes hit masks that are valid for this sweep, 2: neighbour lists; 0: the sweep has
// to filter; < 0: error
int aqc_pc_prepare(aqc_ctx* ctx, const int* imove, const void* r, const void* rj, int dims, float cut2,
                   const LLParams& ll, uint32_t icls, uint32_t jcls, int K, S3Cache* out)
{
    // rj != nullptr: the cache of the remote (halo) sweeps (i cells from ll.icell_i, j rows at rj)
    const bool remote = rj != nullptr;
    aqc_pair_cache& c = remote ? ctx->pcr : ctx->pc;
    c.remote = remote;
    if (!c.enabled || ((icls | jcls) & ~31u) || (!remote && ll.icell_i != ll.icell) || ll.cls)
        return 0;
    {
        static int lists_env = -1;
        if (lists_env < 0) {
            const char* e = getenv("AQC_PAIR_LISTS");
            lists_env = (e && atoi(e) == 0) ? 0 : 1;
        }
        if (!c.builds)
            c.lists = lists_env == 1;
    }
    // a list holds the pairs that passed the exact test and nothing else: a reader whose j set
    // differs from the one the lists were made for cannot use them (the masks can, by re-testing)
    if (c.lists && c.jcls_want && jcls != c.jcls_want)
        return 0;
    const bool same = c.valid && c.r == r && c.rj == rj && c.icell_i == ll.icell_i && c.imove == imove &&
                      c.icell == ll.icell && c.ihoc == ll.ihoc &&
                      c.N == ll.N && c.nx == ll.nx && c.ny == ll.ny && c.nz == ll.nz && c.nw == ll.nw &&
                      c.dims == dims && c.cut2 == cut2 && !(icls & ~c.icls) && !(jcls & ~c.jcls);
    if (!same) {
        aqc_lanes_drain_other(ctx); // (a reader of the lists on the other lane)
        if (c.cooldown) { // recent builds did not pay
            c.cooldown--;
            c.valid = false;
            return 0;
        }
        if (c.builds) {
            c.poor_streak = (c.served < 2) ? c.poor_streak + 1 : 0;
            if (c.poor_streak >= 3) {
                c.poor_streak = 0;
                c.cooldown = 64;
                c.valid = false;
                return 0;
            }
        }
        c.served = 0;
        c.valid = false;
        c.icls_want |= icls;
        c.jcls_want |= jcls;
        c.r = r; c.imove = imove; c.icell = ll.icell; c.ihoc = ll.ihoc;
        c.rj = rj; c.icell_i = ll.icell_i;
        c.N = ll.N; c.nx = ll.nx; c.ny = ll.ny; c.nz = ll.nz; c.nw = ll.nw;
        c.dims = dims; c.cut2 = cut2;
        const size_t nblk = aqc_blocks(ll.N, S3_PARTICLES);
        if (!c.ctl) {
            AQC_CUDA(ctx, cudaMalloc(&c.ctl, 4 * sizeof(unsigned long long)));
            AQC_CUDA(ctx, cudaMallocHost(&c.ctl_host, 4 * sizeof(unsigned long long)));
        }
        if (nblk * S3_MAXPASS > c.pass_cap) {
            if (c.pass_tab)
                AQC_CUDA(ctx, cudaFree(c.pass_tab));
            c.pass_tab = nullptr;
            c.pass_cap = 0;
            AQC_CUDA(ctx, cudaMalloc(&c.pass_tab, nblk * S3_MAXPASS * sizeof(uint32_t)));
            c.pass_cap = nblk * S3_MAXPASS;
        }
        // bytes of one round: hit masks of every (tile, lane), or one chunk count per lane
        const size_t round_bytes = c.lists ? (size_t)S3_CWARPS * 32 : (size_t)(S3_TILES * S3_CWARPS * 32) * sizeof(uint32_t);
        void** rounds_buf = c.lists ? (void**)&c.cnt : (void**)&c.masks;
        auto no_room = [&]() { // no room next to the problem: the sweeps keep filtering
            (void)cudaGetLastError();
            c.cooldown = 0xFFFFFFFFu;
            return 0;
        };
        size_t want = c.cap_rounds ? c.cap_rounds : nblk * (dims == 3 ? 40 : 12);
        // (a halo list holds at most the far half of a neighbourhood: half the chunks to begin with)
        uint32_t want_capc = c.capc ? c.capc : (dims == 3 ? (remote ? 136u : 272u) : (remote ? 36u : 72u)); // (even)
        for (int attempt = 0;; attempt++) {
            if (want > c.cap_rounds) {
                if (*rounds_buf) {
                    AQC_SYNC(ctx);
                    AQC_CUDA(ctx, cudaFree(*rounds_buf));
                }
                *rounds_buf = nullptr;
                c.cap_rounds = 0;
                if (want >= 0xFFFFFFF0ull)
                    return 0; // beyond the 32-bit round index: no cache
                if (cudaMalloc(rounds_buf, want * round_bytes) != cudaSuccess) {
                    *rounds_buf = nullptr;
                    return no_room();
                }
                c.cap_rounds = want;
            }
            if (c.lists) {
                // (slack: the readers load two chunk pairs ahead and prefetch three more)
                const size_t need = (nblk * S3_CWARPS * (size_t)want_capc + 16) * 32 * sizeof(uint2);
                if (need > c.chunks_bytes || want_capc != c.capc) {
                    if (need > c.chunks_bytes) {
                        if (c.chunks) {
                            AQC_SYNC(ctx);
                            AQC_CUDA(ctx, cudaFree(c.chunks));
                        }
                        c.chunks = nullptr;
                        c.chunks_bytes = 0;
                        if (cudaMalloc(&c.chunks, need) != cudaSuccess) {
                            c.chunks = nullptr;
                            return no_room();
                        }
                        c.chunks_bytes = need;
                    }
                    c.capc = want_capc;
                }
            }
            AQC_CUDA(ctx, cudaMemsetAsync(c.ctl, 0, 4 * sizeof(unsigned long long), ctx->stream));
            AQC_CUDA(ctx, cudaMemsetAsync(c.pass_tab, 0xFF, nblk * S3_MAXPASS * sizeof(uint32_t), ctx->stream));
            const int rc = (dims == 3) ? pc_build<3>(ctx, c, ll, K) : pc_build<2>(ctx, c, ll, K);
            if (rc)
                return rc;
            AQC_CUDA(ctx, cudaMemcpyAsync(c.ctl_host, c.ctl, 4 * sizeof(unsigned long long),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            AQC_SYNC(ctx);
            const bool rounds_ok = c.ctl_host[0] <= c.cap_rounds;
            const bool lists_ok = !c.lists || c.ctl_host[2] <= c.capc;
            if (rounds_ok && lists_ok)
                break;
            if (attempt >= 2)
                return aqc_fail(ctx, AQC_ERR_CUDA, "pair cache: %llu rounds / %llu chunks per lane do not fit "
                                "%zu / %u after growing", c.ctl_host[0], c.ctl_host[2], c.cap_rounds, c.capc);
            if (!rounds_ok)
                want = (size_t)(c.ctl_host[0] + c.ctl_host[0] / 8 + 64);
            if (!lists_ok)
                want_capc = ((uint32_t)(c.ctl_host[2] + c.ctl_host[2] / 8 + 8) + 1u) & ~1u;
        }
        c.unusable = (c.ctl_host[1] & 1ull) != 0;
        c.icls = c.icls_want;
        c.jcls = c.jcls_want;
        c.valid = true;
        c.builds++;
    }
    if (c.unusable)
        return 0;
    c.hits++;
    c.served++;
    out->masks = c.masks;
    out->pass_tab = c.pass_tab;
    out->ctl = c.ctl;
    out->cap_rounds = (uint32_t)c.cap_rounds;
    out->chunks = (uint2*)c.chunks;
    out->cnt = c.cnt;
    out->capc = c.capc;
    return c.lists ? 2 : 1;
}

namespace {

// ------------------------------------------------------------------------
// basic/neighs.cl:52-91: candidates are counted without any distance test, so
// the count is the clamped sum of the neighbour cells' populations.
template <int D>
__global__ void __launch_bounds__(256)
neighs_kernel(const int* __restrict__ imove, uint32_t* __restrict__ n_neighs, uint32_t limit,
              const LLParams ll)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ll.N)
        return;
    if (imove[i] <= -255) {
        n_neighs[i] = 0;
        return;
    }
    // The three cells c-1, c, c+1 of an x row hold one contiguous run of the sorted list:
    // its length comes from two searches instead of a walk over every candidate.
    const uint32_t c = __ldg(ll.icell + i);
    constexpr int KZ = (D == 3) ? 1 : 0;
    uint32_t n = 0;
    for (int cj = -1; cj <= 1; cj++)
        for (int ck = -KZ; ck <= KZ; ck++) {
            const uint32_t lo = c - 1u + (uint32_t)cj * ll.nx + (uint32_t)ck * ll.nx * ll.ny;
            const uint32_t b = min(min(__ldg(ll.ihoc + lo), __ldg(ll.ihoc + lo + 1u)),
                                   __ldg(ll.ihoc + lo + 2u));
            if (b < ll.N)
                n += s3_run_end(ll.icell, b, ll.N, lo, 2u) - b;
        }
    // the reference counts one by one and stops as soon as the count reaches the limit
    if (n >= limit && n)
        n = max(limit, 1u);
    n_neighs[i] = n;
}

// ---- launch glue -------------------------------------------------------------
LLParams make_ll(void* const* a, int k_icell, size_t N)
{
    // LINKLIST_LOCAL_PARAMS = icell, ihoc, n_cells (types.h:106-109)
    LLParams ll;
    ll.icell_i = ll.icell = (const uint32_t*)a[k_icell];
    ll.ihoc = (const uint32_t*)a[k_icell + 1];
    const aqc_u4 nc = aqc_scalar<aqc_u4>(a, k_icell + 2);
    ll.nx = nc.x; ll.ny = nc.y; ll.nz = nc.z; ll.nw = nc.w;
    ll.N = (uint32_t)N;
    return ll;
}

template <class P>
void set_base(P& p, const aqc_ctx* ctx, const void* imove)
{
    p.imove = (const int*)imove;
    p.invH = 1.f / ctx->defs.H;
    const float s = ctx->defs.SUPPORT * ctx->defs.H;
    p.cut2 = s * s;
}

#define DIMS_DISPATCH(ctx, FN, ...)                                            \
    ((ctx)->defs.dims == 3 ? FN<3>(__VA_ARGS__) : FN<2>(__VA_ARGS__))

template <int D, class P> int run_interactions(aqc_ctx* ctx, size_t n, void* const* a)
{
    P p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.u = a[2]; p.rho = (const float*)a[3]; p.m = (const float*)a[4];
    p.p = (const float*)a[5]; p.grad_p = a[6]; p.lap_u = a[7]; p.div_u = (float*)a[8];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    const uint32_t N = aqc_scalar<uint32_t>(a, 9);
    (void)n;
    return launch_sweep(ctx, p, make_ll(a, 10, N));
}
int l_interactions(aqc_ctx* c, size_t n, void* const* a)
{
    if (c->lap_morris) // <Define name="__LAP_FORMULATION__" value="__LAP_MORRIS__"/>
        return c->defs.dims == 3 ? run_interactions<3, PInteractionsMorris<3>>(c, n, a)
                                 : run_interactions<2, PInteractionsMorris<2>>(c, n, a);
    return c->defs.dims == 3 ? run_interactions<3, PInteractions<3>>(c, n, a)
                             : run_interactions<2, PInteractions<2>>(c, n, a);
}

template <int D, int MODE> int run_shepard(aqc_ctx* ctx, void* const* a)
{
    PShepard<D, MODE> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.rho = (const float*)a[2]; p.m = (const float*)a[3]; p.shepard = (float*)a[4];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 6, aqc_scalar<uint32_t>(a, 5)));
}
int l_shepard_basic(aqc_ctx* c, size_t, void* const* a)
{
    return c->defs.dims == 3 ? run_shepard<3, 0>(c, a) : run_shepard<2, 0>(c, a);
}
int l_shepard_cfd(aqc_ctx* c, size_t, void* const* a)
{
    return c->defs.dims == 3 ? run_shepard<3, 1>(c, a) : run_shepard<2, 1>(c, a);
}

// cfd/Boundary/Portal/Shepard.cl: (imove, imirrored, r, rho, m, shepard, N, icell, ihoc, n_cells)
template <int D> int run_portal_shepard(aqc_ctx* ctx, void* const* a)
{
    PPortalShepard<D> p;
    set_base(p, ctx, a[1]); // (the engine's i filter reads imirrored)
    p.mv = (const int*)a[0];
    p.r = a[2]; p.rho = (const float*)a[3]; p.m = (const float*)a[4]; p.shepard = (float*)a[5];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 7, aqc_scalar<uint32_t>(a, 6)));
}
int l_portal_shepard(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_portal_shepard, c, a); }
// cfd/Boundary/Portal/Interactions.cl: (imove, imirrored, r, u, rho, m, p, grad_p, lap_u, div_u, N, icell, ihoc, n_cells)
template <int D, bool MORRIS> int run_portal_inter(aqc_ctx* ctx, void* const* a)
{
    PPortalInteractions<D, MORRIS> p;
    set_base(p, ctx, a[1]);
    p.mv = (const int*)a[0];
    p.r = a[2]; p.u = a[3]; p.rho = (const float*)a[4]; p.m = (const float*)a[5]; p.p = (const float*)a[6];
    p.grad_p = a[7]; p.lap_u = a[8]; p.div_u = (float*)a[9];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, make_ll(a, 11, aqc_scalar<uint32_t>(a, 10)));
}
int l_portal_inter(aqc_ctx* c, size_t, void* const* a)
{
    if (c->lap_morris)
        return c->defs.dims == 3 ? run_portal_inter<3, true>(c, a) : run_portal_inter<2, true>(c, a);
    return c->defs.dims == 3 ? run_portal_inter<3, false>(c, a) : run_portal_inter<2, false>(c, a);
}

template <int D, bool V> int run_deltagrad(aqc_ctx* ctx, void* const* a)
{
    PDeltaGrad<D, V> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.rho = (const float*)a[2]; p.m = (const float*)a[3]; p.p = (const float*)a[4];
    p.out = a[5];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll(a, 7, aqc_scalar<uint32_t>(a, 6)));
}
int l_dsph_full(aqc_ctx* c, size_t, void* const* a)
{
    return c->defs.dims == 3 ? run_deltagrad<3, true>(c, a) : run_deltagrad<2, true>(c, a);
}
int l_dsph_lapp(aqc_ctx* c, size_t, void* const* a)
{
    return c->defs.dims == 3 ? run_deltagrad<3, false>(c, a) : run_deltagrad<2, false>(c, a);
}

template <int D> int run_lapp_corr(aqc_ctx* ctx, void* const* a)
{
    PLappCorr<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.rho = (const float*)a[2]; p.m = (const float*)a[3]; p.lap_p_corr = a[4];
    p.lap_p = (float*)a[5];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll(a, 7, aqc_scalar<uint32_t>(a, 6)));
}
int l_dsph_lapp_corr(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_lapp_corr, c, a); }

template <int D> int run_mls(aqc_ctx* ctx, void* const* a)
{
    PMLS<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.rho = (const float*)a[2]; p.m = (const float*)a[3]; p.mls = (float*)a[4];
    p.mls_imove = aqc_scalar<uint32_t>(a, 6);
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll(a, 7, aqc_scalar<uint32_t>(a, 5)));
}
int l_mls(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_mls, c, a); }

template <int D> int run_sensors(aqc_ctx* ctx, void* const* a)
{
    PSensors<D> p;
    set_base(p, ctx, a[1]);
    p.r = a[2]; p.m = (const float*)a[3]; p.u = a[4]; p.rho = (float*)a[5]; p.p = (float*)a[6];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 9, aqc_scalar<uint32_t>(a, 7)));
}
int l_sensors(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_sensors, c, a); }

template <int D> int run_bie_inter(aqc_ctx* ctx, void* const* a)
{
    PBIeInteractions<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.normal = a[2]; p.u = a[3]; p.m = (const float*)a[4]; p.grad_w_bi = a[5];
    p.div_u_bi = (float*)a[6];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 8, aqc_scalar<uint32_t>(a, 7)));
}
int l_bie_inter(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bie_inter, c, a); }

template <int D> int run_bie_pb(aqc_ctx* ctx, void* const* a)
{
    PBIePBoundary<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.m = (const float*)a[2]; p.rho = (const float*)a[3]; p.p = (float*)a[4];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 6, aqc_scalar<uint32_t>(a, 5)));
}
int l_bie_pb(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bie_pb, c, a); }

template <int D> int run_bie_eb(aqc_ctx* ctx, void* const* a)
{
    PBIeElasticBounce<D> p;
    set_base(p, ctx, a[0]);
    p.r_in = a[1]; p.normal = a[2]; p.m = (const float*)a[3]; p.u_in = a[4]; p.dudt = a[5];
    p.dt = aqc_scalar<float>(a, 7);
    p.dr_factor = ctx->dr_factor;
    p.min_bound = ctx->min_bound_dist;
    return launch_sweep(ctx, p, make_ll(a, 8, aqc_scalar<uint32_t>(a, 6)));
}
int l_bie_eb(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bie_eb, c, a); }

template <int D> int run_bie_pst(aqc_ctx* ctx, void* const* a)
{
    PBIePST<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.normal = a[2]; p.m = (const float*)a[3]; p.rho = (const float*)a[4];
    p.inv_dims = 1.f / ctx->defs.DIMS;
    p.dr_factor = ctx->dr_factor;
    return launch_sweep(ctx, p, make_ll(a, 6, aqc_scalar<uint32_t>(a, 5)));
}
int l_bie_pst(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bie_pst, c, a); }

int l_neighs(aqc_ctx* ctx, size_t, void* const* a)
{
    const uint32_t limit = aqc_scalar<uint32_t>(a, 2);
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    const LLParams ll = make_ll(a, 4, N);
    if (ctx->defs.dims == 3)
        neighs_kernel<3><<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>((const int*)a[0],
                                                                       (uint32_t*)a[1], limit, ll);
    else
        neighs_kernel<2><<<aqc_blocks(N, 256), 256, 0, ctx->stream>>>((const int*)a[0],
                                                                       (uint32_t*)a[1], limit, ll);
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

template <int D> int run_bi_shepard(aqc_ctx* ctx, void* const* a)
{
    PBIShepard<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.normal = a[2]; p.tangent = a[3]; p.binormal = a[4]; p.m = (const float*)a[5];
    p.shepard = (float*)a[6];
    p.H = ctx->defs.H; p.CONW = ctx->defs.CONW; p.inv_dm1 = 1.f / (ctx->defs.DIMS - 1.f);
    return launch_sweep(ctx, p, make_ll(a, 8, aqc_scalar<uint32_t>(a, 7)));
}
int l_bi_shepard(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bi_shepard, c, a); }
template <int D> int run_bi_lapu(aqc_ctx* ctx, void* const* a)
{
    PBILapU<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.u = a[2]; p.rho = (const float*)a[3]; p.m = (const float*)a[4]; p.lap_u = a[5];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, make_ll(a, 7, aqc_scalar<uint32_t>(a, 6)));
}
int l_bi_lapu(aqc_ctx* c, size_t, void* const* a)
{
    if (c->lap_morris) // (the host runs the script itself under this definition: Kernel::setup)
        return aqc_fail(c, AQC_ERR_ARG, "%s: the hand-written kernel holds the __LAP_MONAGHAN__ branch only", "cfd/Boundary/BI/LapU.cl::freeslip");
    return DIMS_DISPATCH(c, run_bi_lapu, c, a);
}
template <int D> int run_bi_interp(aqc_ctx* ctx, void* const* a)
{
    PBIInterpolation<D> p;
    set_base(p, ctx, a[1]);
    p.r = a[2]; p.m = (const float*)a[3]; p.rho = (const float*)a[4]; p.grad_p = a[5];
    p.p = (float*)a[6];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll(a, 9, aqc_scalar<uint32_t>(a, 8)));
}
int l_bi_interp(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bi_interp, c, a); }
template <int D> int run_bi_inter(aqc_ctx* ctx, void* const* a)
{
    // (iset, imove, r, normal, u, rho, m, p, refd, grad_p, div_u, icell, ihoc, N, n_cells, g)
    PBIInteractions<D> p;
    set_base(p, ctx, a[1]);
    p.r = a[2]; p.normal = a[3]; p.u = a[4]; p.rho = (const float*)a[5]; p.m = (const float*)a[6];
    p.p = (const float*)a[7]; p.grad_p = a[9]; p.div_u = (float*)a[10];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    LLParams ll;
    ll.icell_i = ll.icell = (const uint32_t*)a[11];
    ll.ihoc = (const uint32_t*)a[12];
    const aqc_u4 nc = aqc_scalar<aqc_u4>(a, 14);
    ll.nx = nc.x; ll.ny = nc.y; ll.nz = nc.z; ll.nw = nc.w;
    ll.N = aqc_scalar<uint32_t>(a, 13);
    return launch_sweep(ctx, p, ll);
}
int l_bi_inter(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_bi_inter, c, a); }
template <int D> int run_bi_noslip(aqc_ctx* ctx, void* const* a)
{
    // (iset, imove, r, normal, u, rho, m, lap_u, N, noslip_iset, dr, icell, ihoc, n_cells)
    PBINoSlip<D> p;
    set_base(p, ctx, a[1]);
    p.iset = (const uint32_t*)a[0];
    p.r = a[2]; p.normal = a[3]; p.u = a[4]; p.rho = (const float*)a[5]; p.m = (const float*)a[6];
    p.lap_u = a[7];
    p.noslip_iset = aqc_scalar<uint32_t>(a, 9);
    p.dr = aqc_scalar<float>(a, 10);
    p.cW = Wend<D>::W * ctx->defs.CONW;
    p.H2 = ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, make_ll(a, 11, aqc_scalar<uint32_t>(a, 8)));
}
int l_bi_noslip(aqc_ctx* c, size_t, void* const* a)
{
    if (c->lap_morris) // (the host runs the script itself under this definition: Kernel::setup)
        return aqc_fail(c, AQC_ERR_ARG, "%s: the hand-written kernel holds the __LAP_MONAGHAN__ branch only", "cfd/Boundary/BI/NoSlip.cl::entry");
    return DIMS_DISPATCH(c, run_bi_noslip, c, a);
}
template <int D> int run_ig_riemann(aqc_ctx* ctx, void* const* a)
{
    // (iset, imove, r, u, rho, m, p, grad_p, div_u, work_density, gamma, N, icell, ihoc, n_cells)
    PRiemann<D> p;
    set_base(p, ctx, a[1]);
    p.iset = (const uint32_t*)a[0];
    p.r = a[2]; p.u = a[3]; p.rho = (const float*)a[4]; p.m = (const float*)a[5]; p.p = (const float*)a[6];
    p.grad_p = a[7]; p.div_u = (float*)a[8]; p.work_density = (float*)a[9]; p.gamma = (const float*)a[10];
    p.cF = 2.f / ctx->defs.H * (Wend<D>::F * ctx->defs.CONW);
    return launch_sweep(ctx, p, make_ll(a, 12, aqc_scalar<uint32_t>(a, 11)));
}
int l_ig_riemann(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_ig_riemann, c, a); }
template <int D> int run_elastic_bounce(aqc_ctx* ctx, void* const* a)
{
    PElasticBounce<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.normal = a[2]; p.u = a[3]; p.dudt = a[4];
    const float dr = aqc_scalar<float>(a, 6);
    p.dt = aqc_scalar<float>(a, 7);
    const float R = (ctx->has_dr_factor ? ctx->dr_factor : 1.5f) * dr; // ElasticBounce.cl:31-37
    p.R2 = R * R;
    p.min_dist = (ctx->has_min_bound_dist ? ctx->min_bound_dist : 0.3f) * dr;
    p.one_plus_e = 1.f + ctx->elastic_factor;
    return launch_sweep(ctx, p, make_ll(a, 8, aqc_scalar<uint32_t>(a, 5)));
}
int l_elastic_bounce(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_elastic_bounce, c, a); }

// LINKLIST_REMOTE_PARAMS = icell, mpi_icell, mpi_ihoc, n_cells (types.h:117-122)
LLParams make_ll_remote(void* const* a, int k_icell, size_t N)
{
    LLParams ll;
    ll.icell_i = (const uint32_t*)a[k_icell];
    ll.icell = (const uint32_t*)a[k_icell + 1];
    ll.ihoc = (const uint32_t*)a[k_icell + 2];
    const aqc_u4 nc = aqc_scalar<aqc_u4>(a, k_icell + 3);
    ll.nx = nc.x; ll.ny = nc.y; ll.nz = nc.z; ll.nw = nc.w;
    ll.N = (uint32_t)N;
    return ll;
}
template <int D> int run_mpi_gamma(aqc_ctx* ctx, void* const* a)
{
    PMpiGamma<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.mpi_r = a[4]; p.mpi_rho = (const float*)a[5]; p.mpi_m = (const float*)a[6];
    p.shepard = (float*)a[7];
    p.cW = Wend<D>::W * ctx->defs.CONW;
    return launch_sweep(ctx, p, make_ll_remote(a, 9, aqc_scalar<uint32_t>(a, 8)));
}
int l_mpi_gamma(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_mpi_gamma, c, a); }
template <int D> int run_mpi_inter(aqc_ctx* ctx, void* const* a)
{
    PMpiInteractions<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.u = a[2]; p.rho = (const float*)a[3]; p.p = (const float*)a[4];
    p.mpi_r = a[5]; p.mpi_u = a[6]; p.mpi_rho = (const float*)a[7]; p.mpi_p = (const float*)a[8];
    p.mpi_m = (const float*)a[9]; p.grad_p = a[10]; p.lap_u = a[11]; p.div_u = (float*)a[12];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, make_ll_remote(a, 14, aqc_scalar<uint32_t>(a, 13)));
}
int l_mpi_inter(aqc_ctx* c, size_t, void* const* a)
{
    if (c->lap_morris) // (the host runs the script itself under this definition: Kernel::setup)
        return aqc_fail(c, AQC_ERR_ARG, "%s: the hand-written kernel holds the __LAP_MONAGHAN__ branch only", "cfd/MPI.cl::interactions");
    return DIMS_DISPATCH(c, run_mpi_inter, c, a);
}

// aqua/MPIdeltaSPH.cl (ours): remote terms of MLS and delta-SPH
template <int D> int run_mpi_mls(aqc_ctx* ctx, void* const* a)
{
    // imove r mpi_r mpi_rho mpi_m mls mls_imove N icell mpi_icell mpi_ihoc n_cells
    PMpiMLS<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.mpi_r = a[2]; p.mpi_rho = (const float*)a[3]; p.mpi_m = (const float*)a[4];
    p.mls = (float*)a[5];
    p.mls_imove = aqc_scalar<uint32_t>(a, 6);
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll_remote(a, 8, aqc_scalar<uint32_t>(a, 7)));
}
int l_mpi_mls(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_mpi_mls, c, a); }
template <int D> int run_mpi_delta(aqc_ctx* ctx, void* const* a)
{
    // imove r p mpi_r mpi_rho mpi_m mpi_p lap_p_corr lap_p N icell mpi_icell mpi_ihoc n_cells
    PMpiDelta<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.p = (const float*)a[2]; p.mpi_r = a[3]; p.mpi_rho = (const float*)a[4];
    p.mpi_m = (const float*)a[5]; p.mpi_p = (const float*)a[6]; p.lap_p_corr = a[7];
    p.lap_p = (float*)a[8];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll_remote(a, 10, aqc_scalar<uint32_t>(a, 9)));
}
int l_mpi_delta(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_mpi_delta, c, a); }
template <int D> int run_mpi_lapp_corr(aqc_ctx* ctx, void* const* a)
{
    // imove r lap_p_corr mpi_r mpi_rho mpi_m mpi_lap_p_corr lap_p N icell mpi_icell mpi_ihoc n_cells
    PMpiLappCorr<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.lap_p_corr = a[2]; p.mpi_r = a[3]; p.mpi_rho = (const float*)a[4];
    p.mpi_m = (const float*)a[5]; p.mpi_lap_p_corr = a[6]; p.lap_p = (float*)a[7];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    return launch_sweep(ctx, p, make_ll_remote(a, 9, aqc_scalar<uint32_t>(a, 8)));
}
int l_mpi_lapp_corr(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_mpi_lapp_corr, c, a); }

// ---- fused launches -----------------------------------------------------------
// Members are given in pipeline order; the argument list of a fused launch is the
// concatenation of the members' own argument lists (same names, same order).
struct FusedEntry {
    std::vector<const char*> members; // "script::entry"
    int (*fn)(aqc_ctx*, void* const*);
};
template <int D, bool SHEP, bool FULL, bool LAPP> int run_fused_fluid(aqc_ctx* ctx, void* const* a)
{
    // [Shepard: imove r rho m shepard N icell ihoc n_cells (9)]
    // Interactions: imove r u rho m p grad_p lap_u div_u N icell ihoc n_cells (13)
    // [full: imove r rho m p lap_p_corr N icell ihoc n_cells (10)] [lapp: ... lap_p ... (10)]
    int k = 0;
    PFusedFluid<D, SHEP, FULL, LAPP> p;
    p.shepard = nullptr; p.lap_p_corr = nullptr; p.lap_p = nullptr;
    if (SHEP) {
        p.shepard = (float*)a[k + 4];
        k += 9;
    }
    void* const* ia = a + k;
    set_base(p, ctx, ia[0]);
    p.r = ia[1]; p.u = ia[2]; p.rho = (const float*)ia[3]; p.m = (const float*)ia[4];
    p.p = (const float*)ia[5]; p.grad_p = ia[6]; p.lap_u = ia[7]; p.div_u = (float*)ia[8];
    const uint32_t N = aqc_scalar<uint32_t>(ia, 9);
    const LLParams ll = make_ll(ia, 10, N);
    k += 13;
    if (FULL) {
        p.lap_p_corr = a[k + 5];
        k += 10;
    }
    if (LAPP) {
        p.lap_p = (float*)a[k + 5];
        k += 10;
    }
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.cW = Wend<D>::W * ctx->defs.CONW;
    p.cWF = p.cW / p.cF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, ll);
}
template <bool SHEP, bool FULL, bool LAPP> int l_fused_fluid(aqc_ctx* c, void* const* a)
{
    return c->defs.dims == 3 ? run_fused_fluid<3, SHEP, FULL, LAPP>(c, a)
                             : run_fused_fluid<2, SHEP, FULL, LAPP>(c, a);
}
template <int D, bool DELTA> int run_mpi_fused(aqc_ctx* ctx, void* const* a)
{
    // interactions: imove r u rho p mpi_r mpi_u mpi_rho mpi_p mpi_m grad_p lap_u div_u N
    //               icell mpi_icell mpi_ihoc n_cells (18); gamma: imove r rho m mpi_r mpi_rho
    //               mpi_m shepard N icell mpi_icell mpi_ihoc n_cells (13)
    //               [full_lapp: imove r p mpi_r mpi_rho mpi_m mpi_p lap_p_corr lap_p N icell
    //               mpi_icell mpi_ihoc n_cells (14)]
    void* const* g = a + 18;
    PMpiFused<D, DELTA> p;
    p.lap_p_corr = DELTA ? g[13 + 7] : nullptr;
    p.lap_p = DELTA ? (float*)g[13 + 8] : nullptr;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.u = a[2]; p.rho = (const float*)a[3]; p.p = (const float*)a[4];
    p.mpi_r = a[5]; p.mpi_u = a[6]; p.mpi_rho = (const float*)a[7]; p.mpi_p = (const float*)a[8];
    p.mpi_m = (const float*)a[9]; p.grad_p = a[10]; p.lap_u = a[11]; p.div_u = (float*)a[12];
    p.shepard = (float*)g[7];
    p.cF = Wend<D>::F * ctx->defs.CONF;
    p.cWF = (Wend<D>::W * ctx->defs.CONW) / p.cF;
    p.eps2 = 0.01f * ctx->defs.H * ctx->defs.H;
    return launch_sweep(ctx, p, make_ll_remote(a, 14, aqc_scalar<uint32_t>(a, 13)));
}
template <bool DELTA> int l_mpi_fused(aqc_ctx* c, void* const* a)
{
    return c->defs.dims == 3 ? run_mpi_fused<3, DELTA>(c, a) : run_mpi_fused<2, DELTA>(c, a);
}
#define K_SHEP "cfd/Shepard.cl::entry"
#define K_INTER "cfd/Interactions.cl::entry"
#define K_FULL "cfd/deltaSPH.cl::full"
#define K_LAPP "cfd/deltaSPH.cl::lapp"
const std::vector<FusedEntry>& fused_table()
{
    static const std::vector<FusedEntry> t = {
        { { K_SHEP, K_INTER, K_FULL, K_LAPP }, l_fused_fluid<true, true, true> },
        { { K_SHEP, K_INTER }, l_fused_fluid<true, false, false> },
        { { K_INTER, K_FULL, K_LAPP }, l_fused_fluid<false, true, true> },
        { { "cfd/MPI.cl::interactions", "cfd/MPI.cl::gamma", "aqua/MPIdeltaSPH.cl::full_lapp" },
          l_mpi_fused<true> },
        { { "cfd/MPI.cl::interactions", "cfd/MPI.cl::gamma" }, l_mpi_fused<false> },
    };
    return t;
}

#define IN(n, t) { n, t, AQC_ARG_ARRAY_IN }
#define OUT(n, t) { n, t, AQC_ARG_ARRAY_OUT }
#define RO(n, t) { n, t, AQC_ARG_ARRAY_RO } // non-const in the reference's script, never written
#define SC(n, t) { n, t, AQC_ARG_SCALAR }
#define LL_ARGS IN("icell", "usize*"), IN("ihoc", "usize*"), SC("n_cells", "svec4")

aqc_registrar r_inter("cfd/Interactions.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"),
      IN("m", "float*"), IN("p", "float*"), OUT("grad_p", "vec*"), OUT("lap_u", "vec*"),
      OUT("div_u", "float*"), SC("N", "usize"), LL_ARGS }, l_interactions);
aqc_registrar r_shep_b("basic/Shepard.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      OUT("shepard", "float*"), SC("N", "usize"), LL_ARGS }, l_shepard_basic);
aqc_registrar r_shep_c("cfd/Shepard.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      OUT("shepard", "float*"), SC("N", "usize"), LL_ARGS }, l_shepard_cfd);
aqc_registrar r_portal_shep("cfd/Boundary/Portal/Shepard.cl", "entry", 0,
    { IN("imove", "int*"), IN("imirrored", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      OUT("shepard", "float*"), SC("N", "usize"), LL_ARGS }, l_portal_shepard);
aqc_registrar r_portal_inter("cfd/Boundary/Portal/Interactions.cl", "entry", 0,
    { IN("imove", "int*"), IN("imirrored", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"),
      IN("m", "float*"), IN("p", "float*"), OUT("grad_p", "vec*"), OUT("lap_u", "vec*"), OUT("div_u", "float*"),
      SC("N", "usize"), LL_ARGS }, l_portal_inter);
#define DSPH_GRAD_ARGS(outname, outtype)                                       \
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"), \
      IN("p", "float*"), OUT(outname, outtype), SC("N", "usize"), LL_ARGS }
aqc_registrar r_full_c("cfd/deltaSPH.cl", "full", 0, DSPH_GRAD_ARGS("lap_p_corr", "vec*"), l_dsph_full);
aqc_registrar r_lapp_c("cfd/deltaSPH.cl", "lapp", 0, DSPH_GRAD_ARGS("lap_p", "float*"), l_dsph_lapp);
#define LAPP_CORR_ARGS                                                          \
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"), \
      IN("lap_p_corr", "vec*"), OUT("lap_p", "float*"), SC("N", "usize"), LL_ARGS }
aqc_registrar r_lappc_c("cfd/deltaSPH.cl", "lapp_corr", 0, LAPP_CORR_ARGS, l_dsph_lapp_corr);
aqc_registrar r_mls("basic/MLS.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      OUT("mls", "matrix*"), SC("N", "usize"), SC("mls_imove", "uint"), LL_ARGS }, l_mls);
aqc_registrar r_sensors("cfd/Sensors.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("m", "float*"),
      OUT("u", "vec*"), OUT("rho", "float*"), OUT("p", "float*"), SC("N", "usize"),
      SC("g", "vec"), LL_ARGS }, l_sensors);
aqc_registrar r_bie_i("cfd/Boundary/BIe/Interactions.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"), IN("u", "vec*"),
      IN("m", "float*"), OUT("grad_w_bi", "vec*"), OUT("div_u_bi", "float*"),
      SC("N", "usize"), LL_ARGS }, l_bie_inter);
aqc_registrar r_bie_pb("cfd/Boundary/BIe/Interactions.cl", "p_boundary", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("m", "float*"), IN("rho", "float*"),
      OUT("p", "float*"), SC("N", "usize"), LL_ARGS }, l_bie_pb);
aqc_registrar r_bie_eb("cfd/Boundary/BIe/ElasticBounce.cl", "entry", 0,
    { IN("imove", "int*"), IN("r_in", "vec*"), IN("normal", "vec*"), IN("m", "float*"),
      IN("u_in", "vec*"), OUT("dudt", "vec*"), SC("N", "usize"), SC("dt", "float"), LL_ARGS },
    l_bie_eb);
aqc_registrar r_bie_pst("cfd/Boundary/BIe/PST.cl", "entry", 0,
    { IN("imove", "int*"), OUT("r", "vec*"), IN("normal", "vec*"), IN("m", "float*"),
      IN("rho", "float*"), SC("N", "usize"), LL_ARGS }, l_bie_pst);
aqc_registrar r_bi_shep("cfd/Boundary/BI/Shepard.cl", "compute", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"), IN("tangent", "vec*"),
      IN("binormal", "vec*"), IN("m", "float*"), OUT("shepard", "float*"), SC("N", "usize"),
      LL_ARGS }, l_bi_shepard);
aqc_registrar r_bi_lapu("cfd/Boundary/BI/LapU.cl", "freeslip", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      OUT("lap_u", "vec*"), SC("N", "usize"), LL_ARGS }, l_bi_lapu);
aqc_registrar r_bi_interp("cfd/Boundary/BI/Interpolation.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("m", "float*"),
      IN("rho", "float*"), IN("grad_p", "vec*"), OUT("p", "float*"), IN("refd", "float*"),
      SC("N", "usize"), LL_ARGS }, l_bi_interp);
aqc_registrar r_bi_inter("cfd/Boundary/BI/Interactions.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"),
      IN("u", "vec*"), IN("rho", "float*"), IN("m", "float*"), IN("p", "float*"),
      IN("refd", "float*"), OUT("grad_p", "vec*"), OUT("div_u", "float*"), RO("icell", "uint*"),
      RO("ihoc", "uint*"), SC("N", "usize"), SC("n_cells", "uivec4"), SC("g", "vec") },
    l_bi_inter);
aqc_registrar r_bi_noslip("cfd/Boundary/BI/NoSlip.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"), IN("u", "vec*"),
      IN("rho", "float*"), IN("m", "float*"), OUT("lap_u", "vec*"), SC("N", "usize"),
      SC("noslip_iset", "uint"), SC("dr", "float"), LL_ARGS }, l_bi_noslip);
aqc_registrar r_ig_riemann("cfd/ideal_gas/riemann/Interactions.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"),
      IN("m", "float*"), IN("p", "float*"), OUT("grad_p", "vec*"), OUT("div_u", "float*"),
      OUT("work_density", "float*"), IN("gamma", "float*"), SC("N", "usize"), LL_ARGS }, l_ig_riemann);
aqc_registrar r_elastic("cfd/Boundary/ElasticBounce.cl", "entry", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"), OUT("u", "vec*"),
      OUT("dudt", "vec*"), SC("N", "usize"), SC("dr", "float"), SC("dt", "float"), LL_ARGS },
    l_elastic_bounce);
#define LL_REMOTE_ARGS IN("icell", "usize*"), IN("mpi_icell", "usize*"), IN("mpi_ihoc", "usize*"), SC("n_cells", "svec4")
aqc_registrar r_mpi_gamma("cfd/MPI.cl", "gamma", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("rho", "float*"), IN("m", "float*"),
      IN("mpi_r", "vec*"), IN("mpi_rho", "float*"), IN("mpi_m", "float*"), OUT("shepard", "float*"),
      SC("N", "usize"), LL_REMOTE_ARGS }, l_mpi_gamma);
aqc_registrar r_mpi_inter("cfd/MPI.cl", "interactions", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"), IN("p", "float*"),
      IN("mpi_r", "vec*"), IN("mpi_u", "vec*"), IN("mpi_rho", "float*"), IN("mpi_p", "float*"),
      IN("mpi_m", "float*"), OUT("grad_p", "vec*"), OUT("lap_u", "vec*"), OUT("div_u", "float*"),
      SC("N", "usize"), LL_REMOTE_ARGS }, l_mpi_inter);
aqc_registrar r_mpi_mls("aqua/MPIdeltaSPH.cl", "mls", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("mpi_r", "vec*"), IN("mpi_rho", "float*"),
      IN("mpi_m", "float*"), OUT("mls", "matrix*"), SC("mls_imove", "unsigned int"), SC("N", "usize"),
      LL_REMOTE_ARGS }, l_mpi_mls);
aqc_registrar r_mpi_delta("aqua/MPIdeltaSPH.cl", "full_lapp", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("p", "float*"), IN("mpi_r", "vec*"),
      IN("mpi_rho", "float*"), IN("mpi_m", "float*"), IN("mpi_p", "float*"), OUT("lap_p_corr", "vec*"),
      OUT("lap_p", "float*"), SC("N", "usize"), LL_REMOTE_ARGS }, l_mpi_delta);
aqc_registrar r_mpi_lapp_corr("aqua/MPIdeltaSPH.cl", "lapp_corr", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("lap_p_corr", "vec*"), IN("mpi_r", "vec*"),
      IN("mpi_rho", "float*"), IN("mpi_m", "float*"), IN("mpi_lap_p_corr", "vec*"), OUT("lap_p", "float*"),
      SC("N", "usize"), LL_REMOTE_ARGS }, l_mpi_lapp_corr);
template <int D> int run_count_pairs(aqc_ctx* ctx, void* const* a)
{
    PCountPairs<D> p;
    set_base(p, ctx, a[0]);
    p.r = a[1]; p.n_pairs = (uint32_t*)a[2];
    return launch_sweep(ctx, p, make_ll(a, 4, aqc_scalar<uint32_t>(a, 3)));
}
int l_count_pairs(aqc_ctx* c, size_t, void* const* a) { return DIMS_DISPATCH(c, run_count_pairs, c, a); }
aqc_registrar r_count_pairs("aqua/diag.cl", "count_pairs", 0,
    { IN("imove", "int*"), IN("r", "vec*"), OUT("n_pairs", "uint*"), SC("N", "usize"), LL_ARGS },
    l_count_pairs);

aqc_registrar r_neighs("basic/neighs.cl", "entry", 0,
    { IN("imove", "int*"), OUT("n_neighs", "uint*"), SC("neighs_limit", "uint"),
      SC("N", "usize"), LL_ARGS }, l_neighs);

} // namespace

extern "C" int aqc_fused_lookup(const int* kernel_ids, int n, int dims)
{
    (void)dims;
    if (!kernel_ids || n < 2)
        return AQC_ERR_NOKERNEL;
    const auto& tab = fused_table();
    for (size_t f = 0; f < tab.size(); f++) {
        if ((int)tab[f].members.size() != n)
            continue;
        bool ok = true;
        for (int k = 0; k < n && ok; k++) {
            const char* nm = aqc_kernel_name(kernel_ids[k]);
            ok = nm && !strcmp(nm, tab[f].members[k]);
        }
        if (ok)
            return (int)f;
    }
    return AQC_ERR_NOKERNEL;
}

extern "C" int aqc_fused_prefix(const int* kernel_ids, int n, int dims)
{
    (void)dims;
    if (!kernel_ids || n < 1)
        return 0;
    for (auto& e : fused_table()) {
        if ((int)e.members.size() < n)
            continue;
        bool ok = true;
        for (int k = 0; k < n && ok; k++) {
            const char* nm = aqc_kernel_name(kernel_ids[k]);
            ok = nm && !strcmp(nm, e.members[k]);
        }
        if (ok)
            return 1;
    }
    return 0;
}

// every fused fluid sweep reads, apart from the positions, only rows of fluid particles
extern "C" int aqc_fused_read_rows(int fused_id)
{
    (void)fused_id;
    return AQC_ROWS_FLUID;
}

extern "C" int aqc_kernel_write_rows(int kernel_id)
{
    const char* nm = aqc_kernel_name(kernel_id);
    if (!nm)
        return AQC_ROWS_ANY;
    // cfd/Sensors.cl:57-130 and SensorsRenormalization.cl:42-67 return unless imove == 0
    if (!strcmp(nm, "cfd/Sensors.cl::entry") || !strcmp(nm, "cfd/SensorsRenormalization.cl::entry"))
        return AQC_ROWS_SENSOR;
    // cfd/Boundary/BIe/Interactions.cl:124-137 returns unless imove == -3: p of the boundary elements only
    // (the tuned-liquid-damper presets place it between cfd interactions and the delta-SPH sweeps)
    if (!strcmp(nm, "cfd/Boundary/BIe/Interactions.cl::p_boundary"))
        return AQC_ROWS_BOUNDARY;
    return AQC_ROWS_ANY;
}

// rows of the arrays OTHER than the positions whose values a kernel uses
extern "C" int aqc_kernel_read_rows(int kernel_id)
{
    const char* nm = aqc_kernel_name(kernel_id);
    if (!nm)
        return AQC_ROWS_ANY;
    // cfd/Boundary/BIe/Interactions.cl:63-106: i is a fluid particle (imove == 1) of which only the
    // position is read, j a boundary element (imove == -3): normal, u and m of boundary rows
    if (!strcmp(nm, "cfd/Boundary/BIe/Interactions.cl::entry"))
        return AQC_ROWS_BOUNDARY;
    // ... :136-170: i a boundary element (position only), j a fluid particle: p, m, rho of fluid rows
    if (!strcmp(nm, "cfd/Boundary/BIe/Interactions.cl::p_boundary"))
        return AQC_ROWS_FLUID;
    return AQC_ROWS_ANY;
}

extern "C" int aqc_launch_fused(aqc_ctx* ctx, int fused_id, void* const* args, int nargs)
{
    if (!ctx)
        return AQC_ERR_ARG;
    const auto& tab = fused_table();
    if (fused_id < 0 || fused_id >= (int)tab.size())
        return aqc_fail(ctx, AQC_ERR_NOKERNEL, "aqc_launch_fused: bad id %d", fused_id);
    int want = 0;
    for (auto nm : tab[fused_id].members) {
        const char* sep = strstr(nm, "::");
        const std::string script(nm, sep - nm);
        want += aqc_kernel_nargs(aqc_kernel_lookup(script.c_str(), sep + 2, ctx->defs.dims));
    }
    if (nargs != want)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch_fused: expected %d args, got %d", want, nargs);
    if (ctx->lap_morris)
        return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch_fused: the fused sweeps are built for "
                                          "__LAP_FORMULATION__ = __LAP_MONAGHAN__ only; launch the members one by one");
    for (int k = 0; k < nargs; k++)
        if (!args[k])
            return aqc_fail(ctx, AQC_ERR_ARG, "aqc_launch_fused: argument %d is NULL", k);
    {
        int k0 = 0; // the members' output arrays
        for (auto nm : tab[fused_id].members) {
            const char* sep = strstr(nm, "::");
            const std::string script(nm, sep - nm);
            const int id = aqc_kernel_lookup(script.c_str(), sep + 2, ctx->defs.dims);
            const aqc_arg_info* ai = aqc_kernel_args(id);
            const int na = aqc_kernel_nargs(id);
            for (int k = 0; k < na; k++)
                if (ai[k].kind == AQC_ARG_ARRAY_OUT)
                    aqc_pc_touch(ctx, args[k0 + k], (size_t)ctx->pc.N * aqc_type_bytes(ai[k].type, ctx->defs.dims));
            k0 += na;
        }
    }
    return tab[fused_id].fn(ctx, args);
}

