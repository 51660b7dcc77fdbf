// elementwise.cu -- the per-particle (no neighbour loop) script kernels of the
// hot path, one HBM pass each, with their Kernel-tool registry entries.
// Compiled with -fmad=false: these kernels are bandwidth bound, so keeping the
// reference's operation order without FMA contraction is free and makes them
// bit-exact against the (uncontracted) oracle.
//
// vec = float4 in 3-D (w carried along exactly as the OpenCL float4 arithmetic
// does), float2 in 2-D.
#include <math.h>

#include "aqc_common.cuh"

namespace {

template <int D> struct V;
template <> struct V<3> {
    float4 v;
    __device__ V() {}
    __device__ V(float4 a) : v(a) {}
    __device__ static V splat(float s) { return V(make_float4(s, s, s, s)); }
    __device__ static V ld(const void* p, size_t i) { return V(reinterpret_cast<const float4*>(p)[i]); }
    __device__ void st(void* p, size_t i) const { reinterpret_cast<float4*>(p)[i] = v; }
    __device__ float dot(const V& o) const { return v.x * o.v.x + v.y * o.v.y + v.z * o.v.z + v.w * o.v.w; }
};
template <> struct V<2> {
    float2 v;
    __device__ V() {}
    __device__ V(float2 a) : v(a) {}
    __device__ static V splat(float s) { return V(make_float2(s, s)); }
    __device__ static V ld(const void* p, size_t i) { return V(reinterpret_cast<const float2*>(p)[i]); }
    __device__ void st(void* p, size_t i) const { reinterpret_cast<float2*>(p)[i] = v; }
    __device__ float dot(const V& o) const { return v.x * o.v.x + v.y * o.v.y; }
};
__device__ inline V<3> operator+(V<3> a, V<3> b) { return V<3>(make_float4(a.v.x + b.v.x, a.v.y + b.v.y, a.v.z + b.v.z, a.v.w + b.v.w)); }
__device__ inline V<3> operator-(V<3> a, V<3> b) { return V<3>(make_float4(a.v.x - b.v.x, a.v.y - b.v.y, a.v.z - b.v.z, a.v.w - b.v.w)); }
__device__ inline V<3> operator*(float s, V<3> a) { return V<3>(make_float4(s * a.v.x, s * a.v.y, s * a.v.z, s * a.v.w)); }
__device__ inline V<3> operator/(V<3> a, float s) { return V<3>(make_float4(a.v.x / s, a.v.y / s, a.v.z / s, a.v.w / s)); }
__device__ inline V<3> operator-(V<3> a) { return V<3>(make_float4(-a.v.x, -a.v.y, -a.v.z, -a.v.w)); }
__device__ inline V<2> operator+(V<2> a, V<2> b) { return V<2>(make_float2(a.v.x + b.v.x, a.v.y + b.v.y)); }
__device__ inline V<2> operator-(V<2> a, V<2> b) { return V<2>(make_float2(a.v.x - b.v.x, a.v.y - b.v.y)); }
__device__ inline V<2> operator*(float s, V<2> a) { return V<2>(make_float2(s * a.v.x, s * a.v.y)); }
__device__ inline V<2> operator/(V<2> a, float s) { return V<2>(make_float2(a.v.x / s, a.v.y / s)); }
__device__ inline V<2> operator-(V<2> a) { return V<2>(make_float2(-a.v.x, -a.v.y)); }

template <int D> __device__ V<D> from_f4(aqc_f4 g);
template <> __device__ inline V<3> from_f4<3>(aqc_f4 g) { V<3> r; r.v = make_float4(g.x, g.y, g.z, g.w); return r; }
template <> __device__ inline V<2> from_f4<2>(aqc_f4 g) { V<2> r; r.v = make_float2(g.x, g.y); return r; }

#define GID                                                                     \
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;            \
    if (i >= N)                                                                 \
        return;

#define LAUNCH(ctx, KERNEL, N, ...)                                            \
    do {                                                                       \
        KERNEL<<<aqc_blocks((N), 256), 256, 0, (ctx)->stream>>>(__VA_ARGS__);  \
        AQC_LAUNCH_CHECK(ctx);                                                 \
    } while (0)

#define DISPATCH(ctx, KERNEL, N, ...)                                          \
    do {                                                                       \
        if ((ctx)->defs.dims == 3)                                             \
            LAUNCH(ctx, KERNEL<3>, N, __VA_ARGS__);                            \
        else                                                                   \
            LAUNCH(ctx, KERNEL<2>, N, __VA_ARGS__);                            \
        return AQC_OK;                                                         \
    } while (0)

// ---- basic/time_scheme/euler.cl:65-87 and midpoint.cl:53-75 (same body) -------
template <int D>
__global__ void __launch_bounds__(256)
k_copy_state(const void* r, const void* u, const void* dudt, const float* rho,
             const float* drhodt, void* r_in, void* u_in, void* dudt_in, float* rho_in,
             float* drhodt_in, uint32_t N)
{
    GID;
    V<D>::ld(dudt, i).st(dudt_in, i);
    V<D>::ld(u, i).st(u_in, i);
    V<D>::ld(r, i).st(r_in, i);
    drhodt_in[i] = drhodt[i];
    rho_in[i] = rho[i];
}
int l_copy_state(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 10);
    DISPATCH(c, k_copy_state, N, a[0], a[1], a[2], (const float*)a[3], (const float*)a[4], a[5],
             a[6], a[7], (float*)a[8], (float*)a[9], N);
}

// ---- basic/time_scheme/euler.cl:105-124 ------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_euler_corrector(const int* imove, void* r, void* u, const void* dudt, float* rho,
                  const float* drhodt, uint32_t N, float dt)
{
    GID;
    if (imove[i] <= 0)
        return;
    const V<D> U = V<D>::ld(u, i), A = V<D>::ld(dudt, i);
    (V<D>::ld(r, i) + (dt * U + (0.5f * dt * dt) * A)).st(r, i);
    (U + dt * A).st(u, i);
    rho[i] = rho[i] + dt * drhodt[i];
}
int l_euler_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 7);
    DISPATCH(c, k_euler_corrector, N, (const int*)a[0], a[2], a[3], a[4], (float*)a[5],
             (const float*)a[6], N, aqc_scalar<float>(a, 8));
}

// ---- basic/time_scheme/improved_euler.cl:75-103 ----------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_ie_predictor(const int* imove, const void* r, const void* u, const void* dudt,
               const float* rho, const float* drhodt, void* r_in, void* u_in, void* dudt_in,
               float* rho_in, float* drhodt_in, uint32_t N, float dt)
{
    GID;
    const float DT = (imove[i] <= 0) ? 0.f : dt;
    const V<D> A = V<D>::ld(dudt, i), U = V<D>::ld(u, i), R = V<D>::ld(r, i);
    A.st(dudt_in, i);
    (U + DT * A).st(u_in, i);
    ((R + DT * U) + (0.5f * DT * DT) * A).st(r_in, i);
    const float dr = drhodt[i];
    drhodt_in[i] = dr;
    rho_in[i] = rho[i] + DT * dr;
}
int l_ie_predictor(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 11);
    DISPATCH(c, k_ie_predictor, N, (const int*)a[0], a[1], a[2], a[3], (const float*)a[4],
             (const float*)a[5], a[6], a[7], a[8], (float*)a[9], (float*)a[10], N,
             aqc_scalar<float>(a, 12));
}

// ---- basic/time_scheme/improved_euler.cl:125-147 ---------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_ie_corrector(const int* imove, void* r, void* u, const void* dudt, float* rho,
               const float* drhodt, const void* dudt_in, const float* drhodt_in, uint32_t N,
               float dt)
{
    GID;
    if (imove[i] <= 0)
        return;
    const float DT = 0.5f * dt;
    const V<D> dA = V<D>::ld(dudt, i) - V<D>::ld(dudt_in, i);
    (V<D>::ld(u, i) + DT * dA).st(u, i);
    (V<D>::ld(r, i) + (DT * DT) * dA).st(r, i);
    rho[i] = rho[i] + DT * (drhodt[i] - drhodt_in[i]);
}
int l_ie_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 9);
    DISPATCH(c, k_ie_corrector, N, (const int*)a[0], a[2], a[3], a[4], (float*)a[5],
             (const float*)a[6], a[7], (const float*)a[8], N, aqc_scalar<float>(a, 10));
}

// ---- basic/time_scheme/midpoint.cl:93-111 ----------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_mp_midpoint(const int* imove, const void* u_in, void* u, const void* dudt,
              const float* rho_in, float* rho, const float* drhodt, uint32_t N, float dt)
{
    GID;
    if (imove[i] <= 0)
        return;
    (V<D>::ld(u_in, i) + (0.5f * dt) * V<D>::ld(dudt, i)).st(u, i);
    rho[i] = rho_in[i] + 0.5f * dt * drhodt[i];
}
int l_mp_midpoint(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 7);
    DISPATCH(c, k_mp_midpoint, N, (const int*)a[0], a[1], a[2], a[3], (const float*)a[4],
             (float*)a[5], (const float*)a[6], N, aqc_scalar<float>(a, 8));
}

// ---- basic/time_scheme/midpoint.cl:124-138 ---------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_mp_midpoint_r(const int* imove, const void* r_in, void* r, const void* u, uint32_t N, float dt)
{
    GID;
    if (imove[i] <= 0)
        return;
    (V<D>::ld(r_in, i) + (0.5f * dt) * V<D>::ld(u, i)).st(r, i);
}
int l_mp_midpoint_r(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    DISPATCH(c, k_mp_midpoint_r, N, (const int*)a[0], a[1], a[2], a[3], N, aqc_scalar<float>(a, 5));
}

// ---- basic/time_scheme/midpoint.cl:141-157 ---------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_mp_relax(const int* imove, const void* dudt_in, void* dudt, const float* drhodt_in,
           float* drhodt, uint32_t N, aqc_sv<float> relax)
{
    GID;
    if (imove[i] <= 0)
        return;
    const float f = relax.get(); // inside a recorded loop: the value the loop's scalar program left
    (f * V<D>::ld(dudt_in, i) + (1.f - f) * V<D>::ld(dudt, i)).st(dudt, i);
    drhodt[i] = f * drhodt_in[i] + (1.f - f) * drhodt[i];
}
int l_mp_relax(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    DISPATCH(c, k_mp_relax, N, (const int*)a[0], a[1], a[2], (const float*)a[3], (float*)a[4], N,
             aqc_scalar_sv<float>(c, a, 6));
}

// ---- basic/time_scheme/midpoint.cl:159-184 ---------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_mp_residuals(const int* imove, const float* m, const void* u, const void* dudt_in,
               const void* dudt, const float* rho, const float* p, const float* drhodt_in,
               const float* drhodt, float* res, uint32_t N)
{
    GID;
    if (imove[i] <= 0) {
        res[i] = 0.f;
        return;
    }
    const float rho2 = rho[i] * rho[i];
    res[i] = m[i] * (fabsf(V<D>::ld(u, i).dot(V<D>::ld(dudt, i) - V<D>::ld(dudt_in, i))) +
                     fabsf(p[i] / rho2 * (drhodt[i] - drhodt_in[i])));
}
int l_mp_residuals(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 10);
    DISPATCH(c, k_mp_residuals, N, (const int*)a[0], (const float*)a[1], a[2], a[3], a[4],
             (const float*)a[5], (const float*)a[6], (const float*)a[7], (const float*)a[8],
             (float*)a[9], N);
}

// ---- basic/time_scheme/midpoint.cl:206-227 ---------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_mp_corrector(const int* imove, const void* r_in, void* r, const void* u_in, void* u,
               const void* dudt, const float* rho_in, float* rho, const float* drhodt, uint32_t N,
               float dt)
{
    GID;
    if (imove[i] <= 0)
        return;
    const V<D> U0 = V<D>::ld(u_in, i), A = V<D>::ld(dudt, i);
    ((V<D>::ld(r_in, i) + dt * U0) + (0.5f * dt * dt) * A).st(r, i);
    (U0 + dt * A).st(u, i);
    rho[i] = rho_in[i] + dt * drhodt[i];
}
int l_mp_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 9);
    DISPATCH(c, k_mp_corrector, N, (const int*)a[0], a[1], a[2], a[3], a[4], a[5],
             (const float*)a[6], (float*)a[7], (const float*)a[8], N, aqc_scalar<float>(a, 10));
}

// ---- basic/Domain.cl:48-90 ---------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_domain(int* imove, void* r_in, void* u_in, void* dudt_in, float* m, uint32_t N, aqc_f4 dmin,
         aqc_f4 dmax)
{
    GID;
    if (imove[i] <= -255)
        return;
    const V<D> c = V<D>::ld(r_in, i);
    bool out = isnan(c.v.x) || isinf(c.v.x) || isnan(c.v.y) || isinf(c.v.y) || (c.v.x < dmin.x) ||
               (c.v.y < dmin.y) || (c.v.x > dmax.x) || (c.v.y > dmax.y);
    if constexpr (D == 3)
        out = out || isnan(c.v.z) || isinf(c.v.z) || (c.v.z < dmin.z) || (c.v.z > dmax.z);
    if (!out)
        return;
    imove[i] = -256;
    m[i] = 0.f;
    V<D>::splat(0.f).st(u_in, i);
    V<D>::splat(0.f).st(dudt_in, i);
    from_f4<D>(dmax).st(r_in, i);
}
int l_domain(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    const int d = c->defs.dims;
    DISPATCH(c, k_domain, N, (int*)a[0], a[1], a[2], a[3], (float*)a[4], N,
             aqc_vec_scalar(a, 6, d), aqc_vec_scalar(a, 7, d));
}

// ---- cfd/Motions/{Transform,UnTransform,Velocity,Acceleration}.cl (preset cfd/motion.xml:55-72) --
// Rigid motion of the non-fluid particles of one set, Euler-XYZ angles.  The scripts evaluate
// cos / sin of the three (uniform) angles in every work-item; here the launcher does it once on
// the host and passes the six values, so the rotations are the same fp32 products and sums.
struct MotionTrig { float cphi, sphi, cth, sth, cpsi, spsi; };
static MotionTrig motion_trig(const aqc_f4& a, float sgn)
{
    return MotionTrig{ cosf(a.x), sgn * sinf(a.x), cosf(a.y), sgn * sinf(a.y), cosf(a.z), sgn * sinf(a.z) };
}
__device__ inline void rot_x(float& y, float& z, float c, float s)
{
    const float y0 = y, z0 = z;
    y = c * y0 - s * z0;
    z = s * y0 + c * z0;
}
__device__ inline void rot_y(float& x, float& z, float c, float s)
{
    const float x0 = x, z0 = z;
    x = c * x0 + s * z0;
    z = -s * x0 + c * z0;
}
__device__ inline void rot_z(float& x, float& y, float c, float s)
{
    const float x0 = x, y0 = y;
    x = c * x0 - s * y0;
    y = s * x0 + c * y0;
}
template <int D> __device__ inline void rotate_fwd(V<D>& a, const MotionTrig& t)
{
    if constexpr (D == 3) {
        rot_x(a.v.y, a.v.z, t.cphi, t.sphi);
        rot_y(a.v.x, a.v.z, t.cth, t.sth);
    }
    rot_z(a.v.x, a.v.y, t.cpsi, t.spsi);
}
template <int D> __device__ inline void rotate_back(V<D>& a, const MotionTrig& t) // t holds -sin
{
    rot_z(a.v.x, a.v.y, t.cpsi, t.spsi);
    if constexpr (D == 3) {
        rot_y(a.v.x, a.v.z, t.cth, t.sth);
        rot_x(a.v.y, a.v.z, t.cphi, t.sphi);
    }
}
// Transform.cl:68-134
template <int D>
__global__ void __launch_bounds__(256)
k_motion_transform(const uint32_t* iset, const int* imove, void* r, void* normal, void* tangent, uint32_t N,
                   uint32_t motion_iset, aqc_f4 motion_r, MotionTrig t)
{
    GID;
    if (iset[i] != motion_iset || imove[i] == 1)
        return;
    V<D> r_i = V<D>::ld(r, i), n_i = V<D>::ld(normal, i), t_i = V<D>::ld(tangent, i);
    rotate_fwd<D>(r_i, t);
    rotate_fwd<D>(n_i, t);
    rotate_fwd<D>(t_i, t);
    (r_i + from_f4<D>(motion_r)).st(r, i);
    (n_i / sqrtf(n_i.dot(n_i))).st(normal, i);   // normalize() over the whole vec
    (t_i / sqrtf(t_i.dot(t_i))).st(tangent, i);
}
int l_motion_transform(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    aqc_f4 ang;
    memcpy(&ang, a[8], sizeof(ang));
    DISPATCH(c, k_motion_transform, N, (const uint32_t*)a[0], (const int*)a[1], a[2], a[3], a[4], N,
             aqc_scalar<uint32_t>(a, 6), aqc_vec_scalar(a, 7, c->defs.dims), motion_trig(ang, 1.f));
}
// UnTransform.cl:54-119
template <int D>
__global__ void __launch_bounds__(256)
k_motion_untransform(const uint32_t* iset, const int* imove, void* r, void* normal, void* tangent, uint32_t N,
                     uint32_t motion_iset, aqc_f4 motion_r_in, MotionTrig t)
{
    GID;
    if (iset[i] != motion_iset || imove[i] == 1)
        return;
    V<D> r_i = V<D>::ld(r, i) - from_f4<D>(motion_r_in), n_i = V<D>::ld(normal, i), t_i = V<D>::ld(tangent, i);
    rotate_back<D>(r_i, t);
    rotate_back<D>(n_i, t);
    rotate_back<D>(t_i, t);
    r_i.st(r, i);
    n_i.st(normal, i);
    t_i.st(tangent, i);
}
int l_motion_untransform(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    aqc_f4 ang;
    memcpy(&ang, a[8], sizeof(ang));
    DISPATCH(c, k_motion_untransform, N, (const uint32_t*)a[0], (const int*)a[1], a[2], a[3], a[4], N,
             aqc_scalar<uint32_t>(a, 6), aqc_vec_scalar(a, 7, c->defs.dims), motion_trig(ang, -1.f));
}
// Velocity.cl:74-121 and Acceleration.cl: omega x r in the local frame, rotated, plus the linear part
template <int D>
__global__ void __launch_bounds__(256)
k_motion_rate(const uint32_t* iset, const int* imove, const void* r, void* out, uint32_t N,
              uint32_t motion_iset, aqc_f4 lin, aqc_f4 w, MotionTrig t)
{
    GID;
    if (iset[i] != motion_iset || imove[i] == 1)
        return;
    const V<D> p = V<D>::ld(r, i);
    V<D> v = V<D>::splat(0.f);
    if constexpr (D == 2) {
        v.v.x = -w.z * p.v.y;
        v.v.y = w.z * p.v.x;
    } else { // cross(float4, float4): w = 0
        v.v.x = w.y * p.v.z - w.z * p.v.y;
        v.v.y = w.z * p.v.x - w.x * p.v.z;
        v.v.z = w.x * p.v.y - w.y * p.v.x;
    }
    rotate_fwd<D>(v, t);
    (v + from_f4<D>(lin)).st(out, i);
}
int l_motion_velocity(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    aqc_f4 ang, w;
    memcpy(&ang, a[7], sizeof(ang));
    memcpy(&w, a[8], sizeof(w));
    DISPATCH(c, k_motion_rate, N, (const uint32_t*)a[0], (const int*)a[1], a[2], a[3], N,
             aqc_scalar<uint32_t>(a, 5), aqc_vec_scalar(a, 6, c->defs.dims), w, motion_trig(ang, 1.f));
}
int l_motion_acceleration(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    aqc_f4 ang, w;
    memcpy(&ang, a[8], sizeof(ang));
    memcpy(&w, a[9], sizeof(w));
    DISPATCH(c, k_motion_rate, N, (const uint32_t*)a[0], (const int*)a[1], a[2], a[3], N,
             aqc_scalar<uint32_t>(a, 5), aqc_vec_scalar(a, 7, c->defs.dims), w, motion_trig(ang, 1.f));
}

// ---- cfd/Energy/Energy.cl::power (:59-87) and ::energy (:114-145), preset cfd/energy.xml ---------
template <int D>
__global__ void __launch_bounds__(256)
k_energy_power(float* dekdt, float* depdt, float* decdt, const int* imove, const void* u, const float* rho,
               const float* m, const float* p, const void* dudt, const float* drhodt, uint32_t N, aqc_f4 g)
{
    GID;
    if (imove[i] != 1) {
        dekdt[i] = 0.f;
        depdt[i] = 0.f;
        decdt[i] = 0.f;
        return;
    }
    const V<D> u_i = V<D>::ld(u, i);
    depdt[i] = -m[i] * from_f4<D>(g).dot(u_i);
    dekdt[i] = m[i] * u_i.dot(V<D>::ld(dudt, i));
    decdt[i] = m[i] * p[i] / (rho[i] * rho[i]) * drhodt[i];
}
int l_energy_power(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 10);
    DISPATCH(c, k_energy_power, N, (float*)a[0], (float*)a[1], (float*)a[2], (const int*)a[3], a[4],
             (const float*)a[5], (const float*)a[6], (const float*)a[7], a[8], (const float*)a[9], N,
             aqc_vec_scalar(a, 11, c->defs.dims));
}
template <int D>
__global__ void __launch_bounds__(256)
k_energy_energy(float* ek, float* ep, float* ec, const uint32_t* iset, const int* imove, const void* r,
                const void* u, const float* rho, const float* m, const float* refd, uint32_t N, aqc_f4 g,
                float cs)
{
    GID;
    if (imove[i] != 1) {
        ek[i] = 0.f;
        ep[i] = 0.f;
        ec[i] = 0.f;
        return;
    }
    const V<D> u_i = V<D>::ld(u, i);
    ek[i] = 0.5f * m[i] * u_i.dot(u_i);
    ep[i] = -m[i] * from_f4<D>(g).dot(V<D>::ld(r, i));
    const float rho0 = refd[iset[i]];
    ec[i] = m[i] * cs * cs * (rho0 / rho[i] + logf(rho[i] / rho0) - 1.f);
}
int l_energy_energy(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 10);
    DISPATCH(c, k_energy_energy, N, (float*)a[0], (float*)a[1], (float*)a[2], (const uint32_t*)a[3],
             (const int*)a[4], a[5], a[6], (const float*)a[7], (const float*)a[8], (const float*)a[9], N,
             aqc_vec_scalar(a, 11, c->defs.dims), aqc_scalar<float>(a, 12));
}

// ---- basic/time_scheme/adam_bashforth.cl (preset basic/time_scheme/adams_bashforth.xml) -----------
// ::predictor (:93-113) is k_copy_state.  The four history levels as1..as4 travel as small structs.
struct AB4 { void* du[4]; float* dr[4]; };
struct AB4c { const void* du[4]; const float* dr[4]; };
// ::sort :122-146
template <int D>
__global__ void __launch_bounds__(256)
k_ab_sort(AB4c in, AB4 out, const uint32_t* id_sorted, uint32_t N)
{
    GID;
    const size_t o = id_sorted[i];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        V<D>::ld(in.du[l], i).st(out.du[l], o);
        out.dr[l][o] = in.dr[l][i];
    }
}
int l_ab_sort(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 17);
    AB4c in;
    AB4 out;
    for (int l = 0; l < 4; l++) {
        in.du[l] = a[2 * l];
        out.du[l] = a[2 * l + 1];
        in.dr[l] = (const float*)a[8 + 2 * l];
        out.dr[l] = (float*)a[9 + 2 * l];
    }
    DISPATCH(c, k_ab_sort, N, in, out, (const uint32_t*)a[16], N);
}
// DYDT_1..DYDT_5 (:148-159): the same fp32 products and sums, left to right
__device__ inline float ab_rate(unsigned local_iter, float d0, float d1, float d2, float d3, float d4)
{
    if (local_iter < 1)
        return d0;
    if (local_iter < 2)
        return 1.5f * d0 - 0.5f * d1;
    if (local_iter < 3)
        return 23.f / 12.f * d0 - 4.f / 3.f * d1 + 5.f / 12.f * d2;
    if (local_iter < 4)
        return 55.f / 24.f * d0 - 59.f / 24.f * d1 + 37.f / 24.f * d2 - 3.f / 8.f * d3;
    return 1901.f / 720.f * d0 - 1387.f / 360.f * d1 + 109.f / 30.f * d2 - 637.f / 360.f * d3 +
           251.f / 720.f * d4;
}
// ::corrector :207-255
template <int D>
__global__ void __launch_bounds__(256)
k_ab_corrector(const int* imove, void* r, void* u, const void* dudt, float* rho, const float* drhodt, AB4c as,
               uint32_t N, float dt, unsigned local_iter)
{
    GID;
    if (imove[i] <= 0)
        return;
    constexpr int VS = D == 3 ? 4 : 2;
    float* rr = reinterpret_cast<float*>(r) + VS * i;
    float* uu = reinterpret_cast<float*>(u) + VS * i;
    const float* d0 = reinterpret_cast<const float*>(dudt) + VS * i;
#pragma unroll
    for (int k = 0; k < VS; k++) {
        const float a = ab_rate(local_iter, d0[k], reinterpret_cast<const float*>(as.du[0])[VS * i + k],
                                reinterpret_cast<const float*>(as.du[1])[VS * i + k],
                                reinterpret_cast<const float*>(as.du[2])[VS * i + k],
                                reinterpret_cast<const float*>(as.du[3])[VS * i + k]);
        rr[k] += dt * uu[k] + 0.5f * dt * dt * a;
        uu[k] += dt * a;
    }
    rho[i] += dt * ab_rate(local_iter, drhodt[i], as.dr[0][i], as.dr[1][i], as.dr[2][i], as.dr[3][i]);
}
int l_ab_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 15);
    const unsigned iter = aqc_scalar<uint32_t>(a, 17);
    AB4c as;
    for (int l = 0; l < 4; l++) {
        as.du[l] = a[7 + 2 * l];
        as.dr[l] = (const float*)a[8 + 2 * l];
    }
    DISPATCH(c, k_ab_corrector, N, (const int*)a[0], a[2], a[3], a[4], (float*)a[5], (const float*)a[6], as, N,
             aqc_scalar<float>(a, 16), iter < c->ab_steps ? iter : c->ab_steps);
}
// ::postcorrector :281-309
template <int D>
__global__ void __launch_bounds__(256)
k_ab_postcorrector(AB4c as, const void* dudt, const float* drhodt, AB4 in, uint32_t N)
{
    GID;
    V<D>::ld(dudt, i).st(in.du[0], i);
    in.dr[0][i] = drhodt[i];
#pragma unroll
    for (int l = 1; l < 4; l++) {
        V<D>::ld(as.du[l - 1], i).st(in.du[l], i);
        in.dr[l][i] = as.dr[l - 1][i];
    }
}
int l_ab_postcorrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 16);
    AB4c as;
    AB4 in;
    for (int l = 0; l < 3; l++) {
        as.du[l] = a[2 * l];
        as.dr[l] = (const float*)a[2 * l + 1];
    }
    as.du[3] = nullptr;
    as.dr[3] = nullptr;
    for (int l = 0; l < 4; l++) {
        in.du[l] = a[8 + 2 * l];
        in.dr[l] = (float*)a[9 + 2 * l];
    }
    DISPATCH(c, k_ab_postcorrector, N, as, a[6], (const float*)a[7], in, N);
}

// ---- small presets next to the hot path -----------------------------------------------------------
// cfd/Energy/EnergyKin.cl:38-54 (preset cfd/energy_kin.xml)
template <int D>
__global__ void __launch_bounds__(256)
k_energy_kin(float* energy_kin, const int* imove, const void* u, const float* m, uint32_t N)
{
    GID;
    if (imove[i] != 1) {
        energy_kin[i] = 0.f;
        return;
    }
    const V<D> u_i = V<D>::ld(u, i);
    energy_kin[i] = 0.5f * m[i] * u_i.dot(u_i);
}
int l_energy_kin(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    DISPATCH(c, k_energy_kin, N, (float*)a[0], (const int*)a[1], a[2], (const float*)a[3], N);
}
// cfd/Forces/Forces.cl:50-84 (preset cfd/forces.xml): force and moment of every fluid particle
template <int D>
__global__ void __launch_bounds__(256)
k_forces(void* forces_f, float4* forces_m, const int* imove, const void* r, const void* dudt, const float* m,
         uint32_t N, aqc_f4 g, aqc_f4 forces_r)
{
    GID;
    if (imove[i] != 1) {
        V<D>::splat(0.f).st(forces_f, i);
        forces_m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const V<D> arm = V<D>::ld(r, i) - from_f4<D>(forces_r);
    const V<D> acc = from_f4<D>(g) - V<D>::ld(dudt, i);
    const float mass = m[i];
    (mass * acc).st(forces_f, i);
    float4 mo = make_float4(0.f, 0.f, mass * (arm.v.x * acc.v.y - arm.v.y * acc.v.x), 0.f);
    if constexpr (D == 3) {
        mo.x = mass * (arm.v.y * acc.v.z - arm.v.z * acc.v.y);
        mo.y = mass * (arm.v.z * acc.v.x - arm.v.x * acc.v.z);
    }
    forces_m[i] = mo;
}
int l_forces(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    const int d = c->defs.dims;
    DISPATCH(c, k_forces, N, a[0], (float4*)a[1], (const int*)a[2], a[3], a[4], (const float*)a[5], N,
             aqc_vec_scalar(a, 7, d), aqc_vec_scalar(a, 8, d));
}
// ---- cfd/Boundary/Symmetry/Mirror.cl:37-251 (preset cfd/symmetry.xml): an infinite symmetry plane made of
// mirrored copies of the particles within the kernel support of it, taken from the buffer rows
template <int D> __device__ inline float sym_dot_xyz(V<D> a, V<D> b);
template <> __device__ inline float sym_dot_xyz<3>(V<3> a, V<3> b) { return a.v.x * b.v.x + a.v.y * b.v.y + a.v.z * b.v.z; }
template <> __device__ inline float sym_dot_xyz<2>(V<2> a, V<2> b) { return a.v.x * b.v.x + a.v.y * b.v.y; }
// base + reflection(u, n) over the XYZ components (Mirror.cl:114-117); w of `base` is kept
template <int D> __device__ inline V<D> sym_reflect_add(V<D> base, V<D> u, V<D> n);
template <> __device__ inline V<3> sym_reflect_add<3>(V<3> base, V<3> u, V<3> n)
{
    const float f = -2.f * sym_dot_xyz<3>(u, n);
    return V<3>(make_float4(base.v.x + f * n.v.x, base.v.y + f * n.v.y, base.v.z + f * n.v.z, base.v.w));
}
template <> __device__ inline V<2> sym_reflect_add<2>(V<2> base, V<2> u, V<2> n)
{
    const float f = -2.f * sym_dot_xyz<2>(u, n);
    return V<2>(make_float2(base.v.x + f * n.v.x, base.v.y + f * n.v.y));
}
// ::drop (:37-55)
template <int D>
__global__ void __launch_bounds__(256)
k_sym_drop(int* imove, void* r, uint32_t N, aqc_f4 symmetry_r, aqc_f4 symmetry_n, aqc_f4 domain_max)
{
    GID;
    if (imove[i] <= -255)
        return;
    const float dr_n = (V<D>::ld(r, i) - from_f4<D>(symmetry_r)).dot(from_f4<D>(symmetry_n));
    if (dr_n >= 0.f) {
        V<D> one = V<D>::splat(1.f);
        if constexpr (D == 3)
            one.v.w = 0.f; // VEC_ONE (types/3D.h:38)
        (from_f4<D>(domain_max) + one).st(r, i);
        imove[i] = -256;
    }
}
int l_sym_drop(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    const int d = c->defs.dims;
    DISPATCH(c, k_sym_drop, N, (int*)a[0], a[1], N, aqc_vec_scalar(a, 3, d), aqc_vec_scalar(a, 4, d),
             aqc_vec_scalar(a, 5, d));
}
// ::detect (:72-96)
template <int D>
__global__ void __launch_bounds__(256)
k_sym_detect(const int* imove, const void* r_in, uint32_t* imirror, uint32_t N, aqc_f4 symmetry_r,
             aqc_f4 symmetry_n, float support_h)
{
    GID;
    if (imove[i] <= -255) {
        imirror[i] = 0u;
        return;
    }
    const float dr_n = (from_f4<D>(symmetry_r) - V<D>::ld(r_in, i)).dot(from_f4<D>(symmetry_n));
    imirror[i] = (fabsf(dr_n) <= support_h) ? 1u : 0u;
}
int l_sym_detect(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    const int d = c->defs.dims;
    DISPATCH(c, k_sym_detect, N, (const int*)a[0], a[1], (uint32_t*)a[2], N, aqc_vec_scalar(a, 4, d),
             aqc_vec_scalar(a, 5, d), c->defs.SUPPORT * c->defs.H);
}
// ::feed (:140-186).  A row that is being fed (a buffer row) is never a row that feeds: its sorted
// imirror entry is 0, so reading its imove before or after the copy leads to the same return.
template <int D>
__global__ void __launch_bounds__(256)
k_sym_feed(int* imove, uint32_t* iset, const uint32_t* imirror, const uint32_t* imirror_invperm,
           uint32_t* mirror_src, void* normal, void* tangent, void* r_in, uint32_t N, uint32_t nbuffer,
           aqc_f4 symmetry_r, aqc_f4 symmetry_n)
{
    GID;
    const int mv = imove[i];
    if (mv <= -255)
        return;
    const uint32_t j = imirror_invperm[i];
    if (imirror[j] != 1u)
        return;
    const uint32_t i0 = N - nbuffer;
    const uint32_t ii = i0 + (N - j - 1u);
    if (ii >= N)
        return;
    const V<D> n = from_f4<D>(symmetry_n);
    mirror_src[ii] = (uint32_t)i;
    imove[ii] = mv;
    iset[ii] = iset[i];
    // (only the XYZ components of the twin's rows are written: .XYZ assignments, Mirror.cl:178-185)
    const V<D> nrm = V<D>::ld(normal, i), tng = V<D>::ld(tangent, i), pos = V<D>::ld(r_in, i);
    V<D> o = sym_reflect_add<D>(nrm, nrm, n);
    V<D> keep = V<D>::ld(normal, ii);
    if constexpr (D == 3)
        o.v.w = keep.v.w;
    o.st(normal, ii);
    o = sym_reflect_add<D>(tng, tng, n);
    keep = V<D>::ld(tangent, ii);
    if constexpr (D == 3)
        o.v.w = keep.v.w;
    o.st(tangent, ii);
    o = sym_reflect_add<D>(pos, pos - from_f4<D>(symmetry_r), n);
    keep = V<D>::ld(r_in, ii);
    if constexpr (D == 3)
        o.v.w = keep.v.w;
    o.st(r_in, ii);
}
int l_sym_feed(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    const int d = c->defs.dims;
    DISPATCH(c, k_sym_feed, N, (int*)a[0], (uint32_t*)a[1], (const uint32_t*)a[2], (const uint32_t*)a[3],
             (uint32_t*)a[4], a[5], a[6], a[7], N, aqc_scalar<uint32_t>(a, 9), aqc_vec_scalar(a, 10, d),
             aqc_vec_scalar(a, 11, d));
}
// ::set (:203-228)
template <int D>
__global__ void __launch_bounds__(256)
k_sym_set(const uint32_t* mirror_src, float* m, void* u_in, void* dudt_in, void* dudt, float* rho_in,
          float* drhodt_in, float* drhodt, uint32_t N, aqc_f4 symmetry_n)
{
    GID;
    const uint32_t src = mirror_src[i];
    if (src >= N)
        return;
    const V<D> n = from_f4<D>(symmetry_n);
    m[i] = m[src];
    rho_in[i] = rho_in[src];
    const float dr = drhodt_in[src];
    drhodt_in[i] = dr;
    drhodt[i] = dr;
    const V<D> us = V<D>::ld(u_in, src), as = V<D>::ld(dudt_in, src);
    V<D> o = sym_reflect_add<D>(us, us, n);
    if constexpr (D == 3)
        o.v.w = V<D>::ld(u_in, i).v.w;
    o.st(u_in, i);
    o = sym_reflect_add<D>(as, as, n);
    V<D> o2 = o;
    if constexpr (D == 3) {
        o.v.w = V<D>::ld(dudt_in, i).v.w;
        o2.v.w = V<D>::ld(dudt, i).v.w;
    }
    o.st(dudt_in, i);
    o2.st(dudt, i);
}
int l_sym_set(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    const int d = c->defs.dims;
    DISPATCH(c, k_sym_set, N, (const uint32_t*)a[0], (float*)a[1], a[2], a[3], a[4], (float*)a[5], (float*)a[6],
             (float*)a[7], N, aqc_vec_scalar(a, 10, d));
}
// ::sort (:239-251)
__global__ void __launch_bounds__(256)
k_sym_sort(const uint32_t* mirror_src_in, uint32_t* mirror_src, const uint32_t* id_sorted, uint32_t N)
{
    GID;
    mirror_src[id_sorted[i]] = mirror_src_in[i];
}
int l_sym_sort(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    LAUNCH(c, k_sym_sort, N, (const uint32_t*)a[0], (uint32_t*)a[1], (const uint32_t*)a[2], N);
    return AQC_OK;
}
// basic/DensityClamp.cl:41-52 (preset basic/densityClamp.xml)
__global__ void __launch_bounds__(256)
k_density_clamp(float* rho_in, uint32_t N, float rho_min, float rho_max)
{
    GID;
    float v = rho_in[i];
    if (v < rho_min)
        v = rho_min;
    if (v > rho_max)
        v = rho_max;
    rho_in[i] = v;
}
int l_density_clamp(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 1);
    LAUNCH(c, k_density_clamp, N, (float*)a[0], N, aqc_scalar<float>(a, 2), aqc_scalar<float>(a, 3));
    return AQC_OK;
}
// basic/IdInverse.cl:33-42 (preset basic/id_inverse.xml)
__global__ void __launch_bounds__(256)
k_id_inverse(const uint32_t* id, uint32_t* id_inverse, uint32_t N)
{
    GID;
    id_inverse[id[i]] = (uint32_t)i;
}
int l_id_inverse(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    LAUNCH(c, k_id_inverse, N, (const uint32_t*)a[0], (uint32_t*)a[1], N);
    return AQC_OK;
}

// ---- cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl, Portal/Mirror.cl (presets cfd/inlet.xml, cfd/outlet.xml,
// cfd/portal.xml): the element-wise kernels of the open boundaries.  Vectors keep every stored component
// (w in 3-D), sums run left to right like the scripts write them.
template <int D> __device__ inline V<D> ob_vec_one();
template <> __device__ inline V<3> ob_vec_one<3>() { return V<3>(make_float4(1.f, 1.f, 1.f, 0.f)); } // types/3D.h:38
template <> __device__ inline V<2> ob_vec_one<2>() { return V<2>(make_float2(1.f, 1.f)); }
// Inlet.cl::feed (:64-126): buffer rows become fluid particles on a lattice of the inlet plane
template <int D>
__global__ void __launch_bounds__(256)
k_inlet_feed(int* imove, const uint32_t* iset, void* r, void* u, void* dudt, float* rho, float* drhodt, float* m,
             float* p, const float* refd, uint32_t first, uint32_t N, float cs, float p0, aqc_f4 g,
             float dr, aqc_f4 inlet_r, aqc_f4 inlet_ru, aqc_f4 inlet_rv, uint32_t inlet_Nx, uint32_t inlet_Ny,
             aqc_f4 inlet_n, float inlet_U, aqc_f4 inlet_rFS, float off)
{
    GID; // (N: the rows to feed; first: the first buffer row, N_particles - nbuffer)
    const size_t ii = (size_t)first + i;
    float u_fac, v_fac;
    if constexpr (D == 2) {
        u_fac = ((float)(uint32_t)i + 0.5f) / (float)inlet_Nx;
        v_fac = 0.f;
    } else {
        const uint32_t u_id = (uint32_t)i % inlet_Nx, v_id = (uint32_t)i / inlet_Nx;
        u_fac = ((float)u_id + 0.5f) / (float)inlet_Nx;
        v_fac = ((float)v_id + 0.5f) / (float)inlet_Ny;
    }
    const V<D> n = from_f4<D>(inlet_n);
    const V<D> ri = from_f4<D>(inlet_r) + u_fac * from_f4<D>(inlet_ru) + v_fac * from_f4<D>(inlet_rv) + off * n;
    ri.st(r, ii);
    imove[ii] = 1;
    V<D>::splat(0.f).st(dudt, ii);
    drhodt[ii] = 0.f;
    (inlet_U * n).st(u, ii);
    const float rd = refd[iset[ii]];
    const float ph = rd * from_f4<D>(g).dot(ri - from_f4<D>(inlet_rFS));
    if constexpr (D == 3)
        m[ii] = rd * dr * dr * dr;
    else
        m[ii] = rd * dr * dr;
    rho[ii] = rd + ph / (cs * cs); // (reversed EOS)
    p[ii] = ph + p0;
}
int l_inlet_feed(aqc_ctx* c, size_t, void* const* a)
{
    if (aqc_scalar<int>(a, 25) == 0) // inlet_starving
        return AQC_OK;
    const int d = c->defs.dims;
    const uint32_t N = aqc_scalar<uint32_t>(a, 10), nbuffer = aqc_scalar<uint32_t>(a, 11);
    uint32_t iN[2];
    memcpy(iN, a[20], sizeof(iN)); // svec2
    const uint64_t want = (uint64_t)iN[0] * iN[1];
    const uint32_t count = (uint32_t)(want < nbuffer ? want : nbuffer);
    if (!count)
        return AQC_OK;
    const float dr = aqc_scalar<float>(a, 16);
    const float off = aqc_scalar<float>(a, 24) - c->defs.SUPPORT * c->defs.H - 0.5f * dr;
    DISPATCH(c, k_inlet_feed, count, (int*)a[0], (const uint32_t*)a[1], a[2], a[3], a[4], (float*)a[5], (float*)a[6],
             (float*)a[7], (float*)a[8], (const float*)a[9], N - nbuffer, count, aqc_scalar<float>(a, 13),
             aqc_scalar<float>(a, 14), aqc_vec_scalar(a, 15, d), dr, aqc_vec_scalar(a, 17, d),
             aqc_vec_scalar(a, 18, d), aqc_vec_scalar(a, 19, d), iN[0], iN[1], aqc_vec_scalar(a, 21, d),
             aqc_scalar<float>(a, 22), aqc_vec_scalar(a, 23, d), off);
}
// Inlet.cl::rates (:148-176): what has not crossed the inlet plane yet moves with the inlet
template <int D>
__global__ void __launch_bounds__(256)
k_inlet_rates(const int* imove, const void* r, void* u, void* dudt, float* drhodt, uint32_t N, aqc_f4 inlet_r,
              float inlet_U, aqc_f4 inlet_n)
{
    GID;
    if (imove[i] != 1)
        return;
    const V<D> n = from_f4<D>(inlet_n);
    if ((V<D>::ld(r, i) - from_f4<D>(inlet_r)).dot(n) > 0.f)
        return;
    (inlet_U * n).st(u, i);
    V<D>::splat(0.f).st(dudt, i);
    drhodt[i] = 0.f;
}
int l_inlet_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    const int d = c->defs.dims;
    DISPATCH(c, k_inlet_rates, N, (const int*)a[0], a[1], a[2], a[3], (float*)a[4], N, aqc_vec_scalar(a, 6, d),
             aqc_scalar<float>(a, 7), aqc_vec_scalar(a, 8, d));
}
// Outlet.cl::rates (:53-96): beyond the outlet plane the particles move with it, hydrostatic
template <int D>
__global__ void __launch_bounds__(256)
k_outlet_rates(const int* imove, const uint32_t* iset, const void* r, void* u, float* rho, float* p, void* dudt,
               void* dudt_in, float* drhodt, float* drhodt_in, const float* refd, uint32_t N, float cs, float p0,
               aqc_f4 g, aqc_f4 outlet_r, aqc_f4 outlet_n, float outlet_U, aqc_f4 outlet_rFS)
{
    GID;
    if (imove[i] != 1)
        return;
    const V<D> ri = V<D>::ld(r, i), n = from_f4<D>(outlet_n);
    if ((ri - from_f4<D>(outlet_r)).dot(n) < 0.f)
        return;
    drhodt[i] = 0.f;
    drhodt_in[i] = 0.f;
    V<D>::splat(0.f).st(dudt, i);
    V<D>::splat(0.f).st(dudt_in, i);
    (outlet_U * n).st(u, i);
    const float rd = refd[iset[i]];
    const float ph = rd * from_f4<D>(g).dot(ri - from_f4<D>(outlet_rFS));
    rho[i] = rd + ph / (cs * cs);
    p[i] = ph + p0;
}
int l_outlet_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 11);
    const int d = c->defs.dims;
    DISPATCH(c, k_outlet_rates, N, (const int*)a[0], (const uint32_t*)a[1], a[2], a[3], (float*)a[4], (float*)a[5],
             a[6], a[7], (float*)a[8], (float*)a[9], (const float*)a[10], N, aqc_scalar<float>(a, 12),
             aqc_scalar<float>(a, 13), aqc_vec_scalar(a, 14, d), aqc_vec_scalar(a, 15, d), aqc_vec_scalar(a, 16, d),
             aqc_scalar<float>(a, 17), aqc_vec_scalar(a, 18, d));
}
// Outlet.cl::feed (:108-131): a kernel support beyond the outlet plane the particles go back to the buffer
template <int D>
__global__ void __launch_bounds__(256)
k_outlet_feed(int* imove, void* r_in, uint32_t N, aqc_f4 domain_max, aqc_f4 outlet_r, aqc_f4 outlet_n,
              float support_h)
{
    GID;
    if (imove[i] != 1)
        return;
    const float dist = (V<D>::ld(r_in, i) - from_f4<D>(outlet_r)).dot(from_f4<D>(outlet_n));
    if (dist < 0.f)
        return;
    if (dist > support_h) {
        (from_f4<D>(domain_max) + ob_vec_one<D>()).st(r_in, i);
        imove[i] = -256;
    }
}
int l_outlet_feed(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    const int d = c->defs.dims;
    DISPATCH(c, k_outlet_feed, N, (int*)a[0], a[1], N, aqc_vec_scalar(a, 3, d), aqc_vec_scalar(a, 4, d),
             aqc_vec_scalar(a, 5, d), c->defs.SUPPORT * c->defs.H);
}
// Portal/Mirror.cl::mirror (:70-95; cell() :32-50 = the link-list's hash): what lies within a kernel support
// of the out plane shows up at the in plane, in the cell it falls into there
template <int D>
__global__ void __launch_bounds__(256)
k_portal_mirror(void* r, int* imirrored, uint32_t* icell, uint32_t N, aqc_f4 portal_in_r, aqc_f4 portal_out_r,
                aqc_f4 portal_n, aqc_f4 r_min, uint32_t nx, uint32_t ny, float support_h, float idist)
{
    GID;
    const V<D> r_ij = V<D>::ld(r, i) - from_f4<D>(portal_out_r);
    if (fabsf(r_ij.dot(from_f4<D>(portal_n))) > support_h) {
        imirrored[i] = 0;
        return;
    }
    imirrored[i] = 1;
    const V<D> q = from_f4<D>(portal_in_r) + r_ij;
    q.st(r, i);
    const uint32_t cx = (uint32_t)((q.v.x - r_min.x) * idist) + 3u;
    const uint32_t cy = (uint32_t)((q.v.y - r_min.y) * idist) + 3u;
    if constexpr (D == 3) {
        const uint32_t cz = (uint32_t)((q.v.z - r_min.z) * idist) + 3u;
        icell[i] = cx - 1u + (cy - 1u) * nx + (cz - 1u) * nx * ny;
    } else {
        icell[i] = cx - 1u + (cy - 1u) * nx;
    }
}
int l_portal_mirror(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    const int d = c->defs.dims;
    uint32_t nc[2];
    memcpy(nc, a[8], sizeof(nc)); // n_cells.x, .y of the uivec4
    const float sh = c->defs.SUPPORT * c->defs.H;
    DISPATCH(c, k_portal_mirror, N, a[0], (int*)a[1], (uint32_t*)a[2], N, aqc_vec_scalar(a, 4, d),
             aqc_vec_scalar(a, 5, d), aqc_vec_scalar(a, 6, d), aqc_vec_scalar(a, 7, d), nc[0], nc[1], sh, 1.f / sh);
}
// Portal/Mirror.cl::unmirror (:108-124)
template <int D>
__global__ void __launch_bounds__(256)
k_portal_unmirror(void* r, const int* imirrored, uint32_t N, aqc_f4 portal_in_r, aqc_f4 portal_out_r)
{
    GID;
    if (!imirrored[i])
        return;
    (from_f4<D>(portal_out_r) + (V<D>::ld(r, i) - from_f4<D>(portal_in_r))).st(r, i);
}
int l_portal_unmirror(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    const int d = c->defs.dims;
    DISPATCH(c, k_portal_unmirror, N, a[0], (const int*)a[1], N, aqc_vec_scalar(a, 3, d), aqc_vec_scalar(a, 4, d));
}
// Portal/Mirror.cl::teleport (:136-153): what crossed the out plane re-enters at the in plane
template <int D>
__global__ void __launch_bounds__(256)
k_portal_teleport(void* r, uint32_t N, aqc_f4 portal_in_r, aqc_f4 portal_out_r, aqc_f4 portal_n)
{
    GID;
    const V<D> r_ij = V<D>::ld(r, i) - from_f4<D>(portal_out_r);
    if (r_ij.dot(from_f4<D>(portal_n)) < 0.f)
        return;
    (from_f4<D>(portal_in_r) + r_ij).st(r, i);
}
int l_portal_teleport(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 1);
    const int d = c->defs.dims;
    DISPATCH(c, k_portal_teleport, N, a[0], N, aqc_vec_scalar(a, 2, d), aqc_vec_scalar(a, 3, d),
             aqc_vec_scalar(a, 4, d));
}
// ---- basic/Sort.cl:57-78 (stage1) and :102-124 (stage2) ----------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_sort_stage1(const uint32_t* id_in, uint32_t* id, const uint32_t* iset_in, uint32_t* iset,
              const int* imove_in, int* imove, const void* r_in, void* r, const void* normal_in,
              void* normal, const void* tangent_in, void* tangent, const uint32_t* id_sorted,
              uint32_t N)
{
    GID;
    const uint32_t o = id_sorted[i];
    id[o] = id_in[i];
    iset[o] = iset_in[i];
    imove[o] = imove_in[i];
    V<D>::ld(r_in, i).st(r, o);
    V<D>::ld(normal_in, i).st(normal, o);
    V<D>::ld(tangent_in, i).st(tangent, o);
}
int l_sort_stage1(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 13);
    DISPATCH(c, k_sort_stage1, N, (const uint32_t*)a[0], (uint32_t*)a[1], (const uint32_t*)a[2],
             (uint32_t*)a[3], (const int*)a[4], (int*)a[5], a[6], a[7], a[8], a[9], a[10], a[11],
             (const uint32_t*)a[12], N);
}
template <int D>
__global__ void __launch_bounds__(256)
k_sort_stage2(const float* rho_in, float* rho, const float* m_in, float* m, const void* u_in,
              void* u, const void* dudt, void* dudt_in, const float* drhodt, float* drhodt_in,
              const uint32_t* id_sorted, uint32_t N)
{
    GID;
    const uint32_t o = id_sorted[i];
    rho[o] = rho_in[i];
    m[o] = m_in[i];
    V<D>::ld(u_in, i).st(u, o);
    // the rates travel the other way round (Sort.cl:119-123)
    V<D>::ld(dudt, i).st(dudt_in, o);
    drhodt_in[o] = drhodt[i];
}
int l_sort_stage2(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 11);
    DISPATCH(c, k_sort_stage2, N, (const float*)a[0], (float*)a[1], (const float*)a[2],
             (float*)a[3], a[4], a[5], a[6], a[7], (const float*)a[8], (float*)a[9],
             (const uint32_t*)a[10], N);
}

// ---- basic/EOS.cl:57-73 ------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_eos(const uint32_t* iset, const int* imove, const float* rho, float* p, const float* refd,
      uint32_t N, float cs, float p0)
{
    GID;
    const int mv = imove[i];
    if ((mv <= 0) && (mv != -1))
        return;
    p[i] = p0 + cs * cs * (rho[i] - refd[iset[i]]);
}
int l_eos(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    LAUNCH(c, k_eos, N, (const uint32_t*)a[0], (const int*)a[1], (const float*)a[2], (float*)a[3],
           (const float*)a[4], N, aqc_scalar<float>(a, 6), aqc_scalar<float>(a, 7));
    return AQC_OK;
}

// ---- basic/Binormal.cl:37-53 -------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_binormal(const void* normal, void* tangent, void* binormal, uint32_t N)
{
    GID;
    if constexpr (D == 2) {
        const float2 n = reinterpret_cast<const float2*>(normal)[i];
        reinterpret_cast<float2*>(binormal)[i] = make_float2(0.f, 0.f);
        reinterpret_cast<float2*>(tangent)[i] = make_float2(n.y, -n.x);
    } else {
        const float4 a = reinterpret_cast<const float4*>(normal)[i];
        const float4 b = reinterpret_cast<const float4*>(tangent)[i];
        reinterpret_cast<float4*>(binormal)[i] =
            make_float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.f);
    }
}
int l_binormal(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    DISPATCH(c, k_binormal, N, a[0], a[1], a[2], N);
}

// ---- cfd/Rates.cl:55-77 ------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_rates(const uint32_t* iset, const int* imove, const void* grad_p, const void* lap_u,
        const float* div_u, void* dudt, float* drhodt, const float* visc_dyn, uint32_t N, aqc_f4 g)
{
    GID;
    if (imove[i] != 1)
        return;
    ((-V<D>::ld(grad_p, i) + visc_dyn[iset[i]] * V<D>::ld(lap_u, i)) + from_f4<D>(g)).st(dudt, i);
    drhodt[i] = -div_u[i];
}
int l_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 9);
    DISPATCH(c, k_rates, N, (const uint32_t*)a[0], (const int*)a[1], a[3], a[4],
             (const float*)a[5], a[6], (float*)a[7], (const float*)a[8], N,
             aqc_vec_scalar(a, 10, c->defs.dims));
}

// ---- cfd/TimeStep.cl:56-77 ---------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_timestep(const int* imove, const void* u, float* dt_var, uint32_t N, float dt, float dt_min,
           float courant, float dt_Ma, float h)
{
    GID;
    if (imove[i] <= 0) {
        dt_var[i] = dt;
        return;
    }
    const V<D> U = V<D>::ld(u, i);
    const float dr_max = dt_Ma * h;
    const float dt_u = courant * dr_max / sqrtf(U.dot(U));
    dt_var[i] = fmaxf(fminf(dt, dt_u), dt_min);
}
int l_timestep(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    DISPATCH(c, k_timestep, N, (const int*)a[0], a[1], (float*)a[2], N, aqc_scalar<float>(a, 4),
             aqc_scalar<float>(a, 5), aqc_scalar<float>(a, 6), aqc_scalar<float>(a, 7),
             aqc_scalar<float>(a, 8));
}

// ---- cfd/SensorsRenormalization.cl:42-67 -------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_sensors_renorm(const int* imove, const float* shepard, void* u, float* rho, float* p, uint32_t N)
{
    GID;
    if (imove[i] != 0)
        return;
    float s = shepard[i];
    if (s < 1.0E-6f)
        s = 1.f;
    (V<D>::ld(u, i) / s).st(u, i);
    rho[i] = rho[i] / s;
    p[i] = p[i] / s;
}
int l_sensors_renorm(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    DISPATCH(c, k_sensors_renorm, N, (const int*)a[0], (const float*)a[1], a[2], (float*)a[3],
             (float*)a[4], N);
}

// ---- basic/deltaSPH.cl:57-71 (simple), :160-173 (full_mls), :330-350 (deltaSPH) -----
template <int D>
__global__ void __launch_bounds__(256)
k_dsph_simple(const uint32_t* iset, const int* imove, void* lap_p_corr, const float* refd,
              uint32_t N, aqc_f4 g)
{
    GID;
    if (imove[i] != 1)
        return;
    (refd[iset[i]] * from_f4<D>(g)).st(lap_p_corr, i);
}
int l_dsph_simple(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    DISPATCH(c, k_dsph_simple, N, (const uint32_t*)a[0], (const int*)a[1], a[2],
             (const float*)a[3], N, aqc_vec_scalar(a, 5, c->defs.dims));
}
template <int D>
__global__ void __launch_bounds__(256)
k_dsph_full_mls(const int* imove, const float* mls, void* lap_p_corr, uint32_t N)
{
    GID;
    if (imove[i] != 1)
        return;
    if constexpr (D == 3) {
        const float4* M = reinterpret_cast<const float4*>(mls) + 4 * i;
        const float4 a = M[0], b = M[1], c = M[2];
        const float4 v = reinterpret_cast<const float4*>(lap_p_corr)[i];
        reinterpret_cast<float4*>(lap_p_corr)[i] =
            make_float4(a.x * v.x + a.y * v.y + a.z * v.z, b.x * v.x + b.y * v.y + b.z * v.z,
                        c.x * v.x + c.y * v.y + c.z * v.z, 0.f);
    } else {
        const float4 M = reinterpret_cast<const float4*>(mls)[i];
        const float2 v = reinterpret_cast<const float2*>(lap_p_corr)[i];
        reinterpret_cast<float2*>(lap_p_corr)[i] =
            make_float2(M.x * v.x + M.y * v.y, M.z * v.x + M.w * v.y);
    }
}
int l_dsph_full_mls(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    DISPATCH(c, k_dsph_full_mls, N, (const int*)a[0], (const float*)a[1], a[2], N);
}
__global__ void __launch_bounds__(256)
k_dsph_apply(const uint32_t* iset, const int* imove, const float* rho, const float* lap_p,
             float* drhodt, const float* refd, const float* delta, uint32_t N, float dt)
{
    GID;
    if (imove[i] != 1)
        return;
    const uint32_t s = iset[i];
    const float delta_f = delta[s] * dt * rho[i] / refd[s];
    drhodt[i] = drhodt[i] + delta_f * lap_p[i];
}
int l_dsph_apply(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 7);
    LAUNCH(c, k_dsph_apply, N, (const uint32_t*)a[0], (const int*)a[1], (const float*)a[2],
           (const float*)a[3], (float*)a[4], (const float*)a[5], (const float*)a[6], N,
           aqc_scalar<float>(a, 8));
    return AQC_OK;
}

// ---- basic/MLS.cl:130-143 (mls_inv); types/3D.h:300-341, 2D.h:250-282 ---------------
__device__ inline void mat3_mul(const float* A, const float* B, float* C)
{
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
            C[a * 3 + b] = A[a * 3 + 0] * B[0 * 3 + b] + A[a * 3 + 1] * B[1 * 3 + b] +
                           A[a * 3 + 2] * B[2 * 3 + b];
}
template <int D>
__global__ void __launch_bounds__(256)
k_mls_inv(const int* imove, float* mls, uint32_t N, uint32_t mls_imove)
{
    GID;
    if ((uint32_t)imove[i] != mls_imove)
        return;
    if constexpr (D == 3) {
        float4* Mp = reinterpret_cast<float4*>(mls) + 4 * i;
        const float4 r0 = Mp[0], r1 = Mp[1], r2 = Mp[2];
        const float M[9] = { r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z };
        float T[9], m[9], I[9], R[9];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
                T[a * 3 + b] = M[b * 3 + a];
        mat3_mul(T, M, m);
        const float det = m[0] * (m[4] * m[8] - m[5] * m[7]) + m[1] * (m[5] * m[6] - m[3] * m[8]) +
                          m[2] * (m[3] * m[7] - m[4] * m[6]);
        const float d = 1.f / det;
        if (fabsf(d) > 1.e16f) {
            I[0] = 1.f; I[1] = 0.f; I[2] = 0.f; I[3] = 0.f; I[4] = 1.f; I[5] = 0.f;
            I[6] = 0.f; I[7] = 0.f; I[8] = 1.f;
        } else {
            I[0] = (m[4] * m[8] - m[5] * m[7]) * d;
            I[1] = (m[2] * m[7] - m[1] * m[8]) * d;
            I[2] = (m[1] * m[5] - m[2] * m[4]) * d;
            I[3] = (m[5] * m[6] - m[3] * m[8]) * d;
            I[4] = (m[0] * m[8] - m[2] * m[6]) * d;
            I[5] = (m[2] * m[3] - m[0] * m[5]) * d;
            I[6] = (m[3] * m[7] - m[4] * m[6]) * d;
            I[7] = (m[1] * m[6] - m[0] * m[7]) * d;
            I[8] = (m[0] * m[4] - m[1] * m[3]) * d;
        }
        mat3_mul(I, T, R);
        Mp[0] = make_float4(R[0], R[1], R[2], 0.f);
        Mp[1] = make_float4(R[3], R[4], R[5], 0.f);
        Mp[2] = make_float4(R[6], R[7], R[8], 0.f);
        Mp[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        float4* Mp = reinterpret_cast<float4*>(mls) + i;
        const float4 M = *Mp;
        const float T0 = M.x, T1 = M.z, T2 = M.y, T3 = M.w; // M.s0213
        const float a = T0 * M.x + T1 * M.z, b = T0 * M.y + T1 * M.w, c = T2 * M.x + T3 * M.z,
                    e = T2 * M.y + T3 * M.w;
        const float d = 1.f / (a * e - b * c);
        float I0, I1, I2, I3;
        if (fabsf(d) > 1.e16f) {
            I0 = 1.f; I1 = 0.f; I2 = 0.f; I3 = 1.f;
        } else {
            I0 = e * d; I1 = -b * d; I2 = -c * d; I3 = a * d;
        }
        *Mp = make_float4(I0 * T0 + I1 * T2, I0 * T1 + I1 * T3, I2 * T0 + I3 * T2,
                          I2 * T1 + I3 * T3);
    }
}
int l_mls_inv(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    DISPATCH(c, k_mls_inv, N, (const int*)a[0], (float*)a[1], N, aqc_scalar<uint32_t>(a, 3));
}

// ---- cfd/Boundary/BIe/Rates.cl:44-62, 76-91, 109-135 --------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_bie_rates(const int* imove, const float* rho, const float* p, const void* u,
            const void* grad_w_bi, const float* div_u_bi, void* grad_p, float* div_u, uint32_t N)
{
    GID;
    if (imove[i] != 1)
        return;
    const V<D> G = V<D>::ld(grad_w_bi, i);
    (V<D>::ld(grad_p, i) + (2.f * p[i] / rho[i]) * G).st(grad_p, i);
    div_u[i] = div_u[i] - 2.f * rho[i] * (V<D>::ld(u, i).dot(G) + div_u_bi[i]);
}
int l_bie_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    DISPATCH(c, k_bie_rates, N, (const int*)a[0], (const float*)a[1], (const float*)a[2], a[3],
             a[4], (const float*)a[5], a[6], (float*)a[7], N);
}
__global__ void __launch_bounds__(256)
k_bie_filter_press(const uint32_t* iset, const int* imove, float* p, uint32_t forces_iset,
                   uint32_t N)
{
    GID;
    if (imove[i] != -3)
        return;
    if (iset[i] != forces_iset)
        p[i] = 0.f;
}
int l_bie_filter_press(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_bie_filter_press, N, (const uint32_t*)a[0], (const int*)a[1], (float*)a[2],
           aqc_scalar<uint32_t>(a, 3), N);
    return AQC_OK;
}
template <int D>
__global__ void __launch_bounds__(256)
k_bie_force_press(const int* imove, const void* r, const void* normal, const float* m,
                  const float* p, void* force_p, float4* moment_p, aqc_f4 fr, uint32_t N)
{
    GID;
    if (imove[i] != -3) {
        V<D>::splat(0.f).st(force_p, i);
        moment_p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float pm = p[i] * m[i];
    if constexpr (D == 3) {
        const float4 n = reinterpret_cast<const float4*>(normal)[i];
        const float4 rr = reinterpret_cast<const float4*>(r)[i];
        const float Fx = pm * n.x, Fy = pm * n.y, Fz = pm * n.z;
        const float Rx = rr.x - fr.x, Ry = rr.y - fr.y, Rz = rr.z - fr.z;
        float* f = reinterpret_cast<float*>(force_p) + 4 * i;
        f[0] = Fx; f[1] = Fy; f[2] = Fz; // force_p[i].XYZ = F.XYZ (w untouched)
        moment_p[i] = make_float4(Ry * Fz - Rz * Fy, Rz * Fx - Rx * Fz, Rx * Fy - Ry * Fx, 0.f);
    } else {
        const float2 n = reinterpret_cast<const float2*>(normal)[i];
        const float2 rr = reinterpret_cast<const float2*>(r)[i];
        const float Fx = pm * n.x, Fy = pm * n.y;
        const float Rx = rr.x - fr.x, Ry = rr.y - fr.y;
        reinterpret_cast<float2*>(force_p)[i] = make_float2(Fx, Fy);
        moment_p[i] = make_float4(Ry * 0.f - 0.f * Fy, 0.f * Fx - Rx * 0.f, Rx * Fy - Ry * Fx, 0.f);
    }
}
int l_bie_force_press(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    DISPATCH(c, k_bie_force_press, N, (const int*)a[0], a[1], a[2], (const float*)a[3],
             (const float*)a[4], a[5], (float4*)a[6], aqc_vec_scalar(a, 7, c->defs.dims), N);
}

// ---- cfd/Boundary/BIe/ElasticBounce.cl:168-184 (force_bound) ------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_bie_force_bound(const int* imove, const float* m, const void* pre, const void* post,
                  void* force, uint32_t N)
{
    GID;
    if (imove[i] != 1) {
        V<D>::splat(0.f).st(force, i);
        return;
    }
    ((-m[i]) * (V<D>::ld(post, i) - V<D>::ld(pre, i))).st(force, i);
}
int l_bie_force_bound(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    DISPATCH(c, k_bie_force_bound, N, (const int*)a[0], (const float*)a[1], a[2], a[3], a[4], N);
}

// ---- basic/SetBuffer.cl:37-49 (count), :64-73 (set_imove) ---------------------------
__global__ void __launch_bounds__(256)
k_setbuffer_count(const int* imove, uint32_t* ibuffer, uint32_t N)
{
    GID;
    ibuffer[i] = (imove[i] == -255) ? 1u : 0u;
}
int l_setbuffer_count(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    LAUNCH(c, k_setbuffer_count, N, (const int*)a[0], (uint32_t*)a[1], N);
    return AQC_OK;
}
__global__ void __launch_bounds__(256)
k_setbuffer_set_imove(int* imove, uint32_t N)
{
    GID;
    if (imove[i] == -256)
        imove[i] = -255;
}
int l_setbuffer_set_imove(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 1);
    LAUNCH(c, k_setbuffer_set_imove, N, (int*)a[0], N);
    return AQC_OK;
}

// ---- cfd/MPI.cl:64-91 (copy), :118-157 (append), :173-198 (remove), :212-228
// (backup_r), :248-267 (sort), :282-295 (eos); cfd/MPI/planes.cl:41-61, 80-99 --------
template <int D>
__global__ void __launch_bounds__(256)
k_mpi_copy(uint32_t* mpi_iset, void* mpi_r, void* mpi_u, void* mpi_dudt, float* mpi_rho,
           float* mpi_drhodt, float* mpi_m, const uint32_t* iset, const void* r, const void* u,
           const void* dudt, const float* rho, const float* drhodt, const float* m, uint32_t N)
{
    GID;
    mpi_iset[i] = iset[i];
    V<D>::ld(r, i).st(mpi_r, i);
    V<D>::ld(u, i).st(mpi_u, i);
    V<D>::ld(dudt, i).st(mpi_dudt, i);
    mpi_rho[i] = rho[i];
    mpi_drhodt[i] = drhodt[i];
    mpi_m[i] = m[i];
}
int l_mpi_copy(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 14);
    DISPATCH(c, k_mpi_copy, N, (uint32_t*)a[0], a[1], a[2], a[3], (float*)a[4], (float*)a[5],
             (float*)a[6], (const uint32_t*)a[7], a[8], a[9], a[10], (const float*)a[11],
             (const float*)a[12], (const float*)a[13], N);
}

template <int D>
__global__ void __launch_bounds__(256)
k_mpi_append(uint32_t* iset, void* r, void* u, void* dudt, float* rho, float* drhodt, float* m,
             int* imove, const uint32_t* mpi_local_mask, const uint32_t* mpi_iset, const void* mpi_r,
             const void* mpi_u, const void* mpi_dudt, const float* mpi_rho, const float* mpi_drhodt,
             const float* mpi_m, uint32_t mpi_rank, uint32_t nbuffer, uint32_t N)
{
    GID;
    if (mpi_local_mask[i] == mpi_rank)
        return;
    const size_t o = (size_t)(N - nbuffer) + i; // MPI.cl:148: buffer particles sit at the end
    imove[o] = 1;
    iset[o] = mpi_iset[i];
    V<D>::ld(mpi_r, i).st(r, o);
    V<D>::ld(mpi_u, i).st(u, o);
    V<D>::ld(mpi_dudt, i).st(dudt, o);
    rho[o] = mpi_rho[i];
    drhodt[o] = mpi_drhodt[i];
    m[o] = mpi_m[i];
}
int l_mpi_append(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 18);
    DISPATCH(c, k_mpi_append, N, (uint32_t*)a[0], a[1], a[2], a[3], (float*)a[4], (float*)a[5],
             (float*)a[6], (int*)a[7], (const uint32_t*)a[8], (const uint32_t*)a[9], a[10], a[11],
             a[12], (const float*)a[13], (const float*)a[14], (const float*)a[15],
             aqc_scalar<uint32_t>(a, 16), aqc_scalar<uint32_t>(a, 17), N);
}

template <int D>
__global__ void __launch_bounds__(256)
k_mpi_remove(int* imove, void* r, void* u, void* dudt, float* m, const uint32_t* mpi_local_mask,
             uint32_t mpi_rank, aqc_f4 domain_max, uint32_t N)
{
    GID;
    if (mpi_local_mask[i] == mpi_rank)
        return;
    imove[i] = -256;
    m[i] = 0.f;
    V<D>::splat(0.f).st(u, i);
    V<D>::splat(0.f).st(dudt, i);
    from_f4<D>(domain_max).st(r, i);
}
int l_mpi_remove(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    DISPATCH(c, k_mpi_remove, N, (int*)a[0], a[1], a[2], a[3], (float*)a[4], (const uint32_t*)a[5],
             aqc_scalar<uint32_t>(a, 6), aqc_vec_scalar(a, 7, c->defs.dims), N);
}

template <int D>
__global__ void __launch_bounds__(256)
k_mpi_backup_r(const uint32_t* mpi_neigh_mask, const void* mpi_r, void* mpi_r_in, aqc_f4 r_max,
               uint32_t mpi_rank, uint32_t N, uint32_t n_radix)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_radix)
        return;
    if ((i >= N) || (mpi_neigh_mask[i] == mpi_rank)) {
        from_f4<D>(r_max).st(mpi_r_in, i);
        return;
    }
    V<D>::ld(mpi_r, i).st(mpi_r_in, i);
}
int l_mpi_backup_r(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t n_radix = aqc_scalar<uint32_t>(a, 6);
    DISPATCH(c, k_mpi_backup_r, n_radix, (const uint32_t*)a[0], a[1], a[2],
             aqc_vec_scalar(a, 3, c->defs.dims), aqc_scalar<uint32_t>(a, 4),
             aqc_scalar<uint32_t>(a, 5), n_radix);
}

template <int D>
__global__ void __launch_bounds__(256)
k_mpi_sort(const uint32_t* mpi_iset_in, uint32_t* mpi_iset, const void* mpi_r_in, void* mpi_r,
           const void* mpi_u_in, void* mpi_u, const float* mpi_rho_in, float* mpi_rho,
           const float* mpi_m_in, float* mpi_m, const uint32_t* mpi_id_sorted, uint32_t N)
{
    GID;
    const size_t o = mpi_id_sorted[i];
    mpi_iset[o] = mpi_iset_in[i];
    V<D>::ld(mpi_r_in, i).st(mpi_r, o);
    V<D>::ld(mpi_u_in, i).st(mpi_u, o);
    mpi_rho[o] = mpi_rho_in[i];
    mpi_m[o] = mpi_m_in[i];
}
int l_mpi_sort(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 11);
    DISPATCH(c, k_mpi_sort, N, (const uint32_t*)a[0], (uint32_t*)a[1], a[2], a[3], a[4], a[5],
             (const float*)a[6], (float*)a[7], (const float*)a[8], (float*)a[9],
             (const uint32_t*)a[10], N);
}

__global__ void __launch_bounds__(256)
k_mpi_eos(const uint32_t* mpi_iset, const float* mpi_rho, float* mpi_p, const float* refd,
          uint32_t N, float cs, float p0)
{
    GID;
    mpi_p[i] = p0 + cs * cs * (mpi_rho[i] - refd[mpi_iset[i]]);
}
int l_mpi_eos(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_mpi_eos, N, (const uint32_t*)a[0], (const float*)a[1], (float*)a[2],
           (const float*)a[3], N, aqc_scalar<float>(a, 5), aqc_scalar<float>(a, 6));
    return AQC_OK;
}

// aqua/MPIdeltaSPH.cl::copy_g / ::sort_g (ours): the corrected pressure gradient of delta-SPH
// travels like the other halo fields -- copied into its mpi_ twin, exchanged, and put in the
// order of the halo link-list (cfd/MPI.cl:64-91 and :248-267 do this for r, u, rho, m)
template <int D>
__global__ void __launch_bounds__(256) k_mpi_copy_g(void* mpi_g, const void* g, uint32_t N)
{
    GID;
    V<D>::ld(g, i).st(mpi_g, i);
}
int l_mpi_copy_g(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 2);
    DISPATCH(c, k_mpi_copy_g, N, a[0], a[1], N);
}
template <int D>
__global__ void __launch_bounds__(256)
k_mpi_sort_g(const void* g_in, void* g, const uint32_t* mpi_id_sorted, uint32_t N)
{
    GID;
    V<D>::ld(g_in, i).st(g, mpi_id_sorted[i]);
}
int l_mpi_sort_g(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    DISPATCH(c, k_mpi_sort_g, N, a[0], a[1], (const uint32_t*)a[2], N);
}

// planes.cl: HALO = false -> local_mask (d > 0), true -> neigh_mask (d >= -SUPPORT*H)
template <int D>
__global__ void __launch_bounds__(256)
k_mpi_plane_mask(const int* imove, const void* r, uint32_t* mask, aqc_f4 plane_r, aqc_f4 plane_n,
                 uint32_t proc, uint32_t N, int halo, float reach)
{
    GID;
    if (imove[i] <= 0)
        return;
    const float d = (V<D>::ld(r, i) - from_f4<D>(plane_r)).dot(from_f4<D>(plane_n));
    if (halo ? (d >= -reach) : (d > 0.f))
        mask[i] = proc;
}
int l_mpi_local_mask(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    DISPATCH(c, k_mpi_plane_mask, N, (const int*)a[0], a[1], (uint32_t*)a[2],
             aqc_vec_scalar(a, 3, c->defs.dims), aqc_vec_scalar(a, 4, c->defs.dims),
             aqc_scalar<uint32_t>(a, 5), N, 0, 0.f);
}
int l_mpi_neigh_mask(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    DISPATCH(c, k_mpi_plane_mask, N, (const int*)a[0], a[1], (uint32_t*)a[2],
             aqc_vec_scalar(a, 3, c->defs.dims), aqc_vec_scalar(a, 4, c->defs.dims),
             aqc_scalar<uint32_t>(a, 5), N, 1, c->defs.SUPPORT * c->defs.H);
}

// ---- cfd/Boundary/BI/GradP.cl:55-83, InterpolationShepard.cl:48-76, Shepard.cl:173-195 ----
template <int D>
__global__ void __launch_bounds__(256)
k_bi_gradp(const uint32_t* iset, const int* imove, const float* shepard, const void* lap_u,
           const void* dudt, void* grad_p, const float* visc_dyn, const float* refd, uint32_t N,
           aqc_f4 g)
{
    GID;
    if (imove[i] != -3)
        return;
    float sh = shepard[i];
    if (sh < 1.0e-6f)
        sh = 1.f;
    const uint32_t s = iset[i];
    const float f = visc_dyn[s] / refd[s];
    (((f * V<D>::ld(lap_u, i)) / sh + from_f4<D>(g)) - V<D>::ld(dudt, i)).st(grad_p, i);
}
int l_bi_gradp(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 9);
    DISPATCH(c, k_bi_gradp, N, (const uint32_t*)a[0], (const int*)a[1], (const float*)a[2], a[4], a[5],
             a[6], (const float*)a[7], (const float*)a[8], N, aqc_vec_scalar(a, 10, c->defs.dims));
}

__global__ void __launch_bounds__(256)
k_bi_interp_shepard(const uint32_t* iset, const int* imove, const float* shepard, float* rho,
                    float* p, const float* refd, uint32_t N, float cs, float p0)
{
    GID;
    if (imove[i] != -3)
        return;
    float sh = shepard[i];
    if (sh < 1.0e-6f)
        sh = 1.f;
    const float pi = p[i] / sh;
    p[i] = pi;
    rho[i] = refd[iset[i]] + (pi - p0) / (cs * cs);
}
int l_bi_interp_shepard(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    LAUNCH(c, k_bi_interp_shepard, N, (const uint32_t*)a[0], (const int*)a[1], (const float*)a[2],
           (float*)a[3], (float*)a[4], (const float*)a[5], N, aqc_scalar<float>(a, 7),
           aqc_scalar<float>(a, 8));
    return AQC_OK;
}

template <int D>
__global__ void __launch_bounds__(256)
k_bi_shepard_apply(const int* imove, const float* shepard, void* grad_p, void* lap_u, float* div_u,
                   uint32_t N)
{
    GID;
    if (imove[i] != 1)
        return;
    const float sh = shepard[i];
    (V<D>::ld(grad_p, i) / sh).st(grad_p, i);
    (V<D>::ld(lap_u, i) / sh).st(lap_u, i);
    div_u[i] = div_u[i] / sh;
}
int l_bi_shepard_apply(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    DISPATCH(c, k_bi_shepard_apply, N, (const int*)a[0], (const float*)a[1], a[2], a[3], (float*)a[4], N);
}

// ---- case-local script of the 3-D dam-break example (wave height probes):
// examples/3D/spheric_testcase2_dambreak/src/templates/h_sensor.cl:1-60 ------------
__global__ void __launch_bounds__(256)
k_h_sensor(const int* imove, const float4* r, float* h_sensorz, uint32_t N, float h_sensorx,
           float dr)
{
    GID;
    if (imove[i] <= 0) {
        h_sensorz[i] = 0.f;
        return;
    }
    const float4 ri = r[i];
    const float x = ri.x - h_sensorx;
    if ((fabsf(x) > 2.f * dr) || (fabsf(ri.y) > 2.f * dr)) {
        h_sensorz[i] = 0.f;
        return;
    }
    h_sensorz[i] = ri.z + 0.5f * dr;
}
int l_h_sensor(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    LAUNCH(c, k_h_sensor, N, (const int*)a[0], (const float4*)a[1], (float*)a[2], N,
           aqc_scalar<float>(a, 4), aqc_scalar<float>(a, 5));
    return AQC_OK;
}

// ---- cfd/ideal_gas: the element-wise kernels of the ideal-gas presets (the internal energy next to
// rho / u; examples/2D/shock_*).  One HBM pass each, the reference's operation order (-fmad=false) ----
// cfd/ideal_gas/EOS.cl:56-70 (EXCLUDED_PARTICLE :32-34)
__global__ void __launch_bounds__(256)
k_ig_eos(const uint32_t* iset, const int* imove, const float* rho, const float* eint, float* p,
         const float* gamma, uint32_t N)
{
    GID;
    const int mv = imove[i];
    if ((mv <= 0) && (mv != -1))
        return;
    p[i] = (gamma[iset[i]] - 1.0f) * rho[i] * eint[i];
}
int l_ig_eos(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    LAUNCH(c, k_ig_eos, N, (const uint32_t*)a[0], (const int*)a[1], (const float*)a[2], (const float*)a[3],
           (float*)a[4], (const float*)a[5], N);
    return AQC_OK;
}

// cfd/ideal_gas/Rates.cl:53-68
__global__ void __launch_bounds__(256)
k_ig_rates(const int* imove, const float* rho, const float* p, const float* div_u, float* deintdt, uint32_t N)
{
    GID;
    if (imove[i] != 1)
        return;
    deintdt[i] = -p[i] / (rho[i] * rho[i]) * div_u[i];
}
int l_ig_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 6);
    LAUNCH(c, k_ig_rates, N, (const int*)a[1], (const float*)a[2], (const float*)a[3], (const float*)a[4],
           (float*)a[5], N);
    return AQC_OK;
}

// cfd/ideal_gas/Sort.cl:43-58
__global__ void __launch_bounds__(256)
k_ig_sort(const float* eint_in, float* eint, const float* deintdt, float* deintdt_in, const uint32_t* id_sorted,
          uint32_t N)
{
    GID;
    const uint32_t o = id_sorted[i];
    eint[o] = eint_in[i];
    deintdt_in[o] = deintdt[i];
}
int l_ig_sort(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    LAUNCH(c, k_ig_sort, N, (const float*)a[0], (float*)a[1], (const float*)a[2], (float*)a[3],
           (const uint32_t*)a[4], N);
    return AQC_OK;
}

// cfd/ideal_gas/TimeStep.cl:62-97 (sound_speed.hcl:22-25; length() of a vec takes every component;
// OpenCL min(x, y) = y < x ? y : x, max(x, y) = x < y ? y : x)
template <int D>
__global__ void __launch_bounds__(256)
k_ig_timestep(float* dt_var, const int* imove, const uint32_t* iset, const void* u, const float* rho,
              const float* p, uint32_t N, float dt, float dt_min, float courant, const float* div_u,
              const void* grad_p, const float* gamma, float H)
{
    GID;
    if (imove[i] <= 0) {
        dt_var[i] = dt;
        return;
    }
    const float dxx = H;
    const float s_i = sqrtf(gamma[iset[i]] * p[i] / rho[i]);
    const V<D> G = V<D>::ld(grad_p, i), U = V<D>::ld(u, i);
    const float lg = sqrtf(G.dot(G)), lu = sqrtf(U.dot(U));
    const float a = 4.0f * dxx * div_u[i] / rho[i];
    const float dt_u1 = courant * 0.4f * dxx / sqrtf(a * a + s_i * s_i);
    const float dt_u2 = courant * sqrtf(dxx / lg);
    const float dt_u3 = courant * 0.4f * dxx / sqrtf(lu * lu + s_i * s_i);
    const float m12 = dt_u2 < dt_u1 ? dt_u2 : dt_u1;
    const float dt_u = dt_u3 < m12 ? dt_u3 : m12;
    const float lo = dt_u < dt ? dt_u : dt;
    dt_var[i] = lo < dt_min ? dt_min : lo;
}
int l_ig_timestep(aqc_ctx* c, size_t, void* const* a)
{
    // (dudt, m and the scalar h are arguments of the script that its body never uses)
    const uint32_t N = aqc_scalar<uint32_t>(a, 8);
    DISPATCH(c, k_ig_timestep, N, (float*)a[0], (const int*)a[1], (const uint32_t*)a[2], a[3], (const float*)a[5],
             (const float*)a[6], N, aqc_scalar<float>(a, 9), aqc_scalar<float>(a, 10), aqc_scalar<float>(a, 11),
             (const float*)a[13], a[14], (const float*)a[15], c->defs.H);
}

// cfd/ideal_gas/riemann/Rates.cl:39-53
__global__ void __launch_bounds__(256)
k_ig_riemann_rates(const int* imove, const float* work_density, float* deintdt, uint32_t N)
{
    GID;
    if (imove[i] != 1)
        return;
    deintdt[i] = -work_density[i];
}
int l_ig_riemann_rates(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_riemann_rates, N, (const int*)a[1], (const float*)a[2], (float*)a[3], N);
    return AQC_OK;
}

// cfd/ideal_gas/time_scheme/midpoint.cl: predictor :47-59, midpoint :75-88, relax :101-115, corrector :131-144
__global__ void __launch_bounds__(256)
k_ig_mp_predictor(const float* eint, const float* deintdt, float* eint_in, float* deintdt_in, uint32_t N)
{
    GID;
    deintdt_in[i] = deintdt[i];
    eint_in[i] = eint[i];
}
int l_ig_mp_predictor(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_mp_predictor, N, (const float*)a[0], (const float*)a[1], (float*)a[2], (float*)a[3], N);
    return AQC_OK;
}
// midpoint (f = 0.5) and corrector (f = 1): 0.5f * dt is exact, so f * dt * rate == the script's expression
__global__ void __launch_bounds__(256)
k_ig_mp_advance(const int* imove, const float* eint_in, const float* deintdt, float* eint, uint32_t N, float fdt)
{
    GID;
    if (imove[i] <= 0)
        return;
    eint[i] = eint_in[i] + fdt * deintdt[i];
}
int l_ig_mp_midpoint(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_mp_advance, N, (const int*)a[0], (const float*)a[1], (const float*)a[2], (float*)a[3], N,
           0.5f * aqc_scalar<float>(a, 5));
    return AQC_OK;
}
int l_ig_mp_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_mp_advance, N, (const int*)a[0], (const float*)a[1], (const float*)a[2], (float*)a[3], N,
           aqc_scalar<float>(a, 5));
    return AQC_OK;
}
__global__ void __launch_bounds__(256)
k_ig_mp_relax(const int* imove, const float* deintdt_in, float* deintdt, uint32_t N, aqc_sv<float> relax)
{
    GID;
    if (imove[i] <= 0)
        return;
    const float f = relax.get(); // (inside a recorded loop: the value the loop's scalar program left)
    deintdt[i] = f * deintdt_in[i] + (1.f - f) * deintdt[i];
}
int l_ig_mp_relax(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    LAUNCH(c, k_ig_mp_relax, N, (const int*)a[0], (const float*)a[1], (float*)a[2], N,
           aqc_scalar_sv<float>(c, a, 4));
    return AQC_OK;
}

// cfd/ideal_gas/time_scheme/euler.cl:70-83 (its predictor :44-56 is k_ig_mp_predictor), improved_euler.cl:49-97
__global__ void __launch_bounds__(256)
k_ig_euler_corrector(const int* imove, float* eint, const float* deintdt, uint32_t N, float dt)
{
    GID;
    if (imove[i] > 0)
        eint[i] += dt * deintdt[i];
}
int l_ig_euler_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 3);
    LAUNCH(c, k_ig_euler_corrector, N, (const int*)a[0], (float*)a[1], (const float*)a[2], N,
           aqc_scalar<float>(a, 4));
    return AQC_OK;
}
__global__ void __launch_bounds__(256)
k_ig_ie_predictor(const int* imove, const float* eint, const float* deintdt, float* eint_in, float* deintdt_in,
                  uint32_t N, float dt)
{
    GID;
    const float DT = (imove[i] <= 0) ? 0.f : dt;
    deintdt_in[i] = deintdt[i];
    eint_in[i] = eint[i] + DT * deintdt[i];
}
int l_ig_ie_predictor(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 5);
    LAUNCH(c, k_ig_ie_predictor, N, (const int*)a[0], (const float*)a[1], (const float*)a[2], (float*)a[3],
           (float*)a[4], N, aqc_scalar<float>(a, 6));
    return AQC_OK;
}
__global__ void __launch_bounds__(256)
k_ig_ie_corrector(const int* imove, const float* deintdt, const float* deintdt_in, float* eint, uint32_t N,
                  float dt)
{
    GID;
    if (imove[i] > 0) {
        const float DT = 0.5f * dt;
        eint[i] += DT * (deintdt[i] - deintdt_in[i]);
    }
}
int l_ig_ie_corrector(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_ie_corrector, N, (const int*)a[0], (const float*)a[1], (const float*)a[2], (float*)a[3], N,
           aqc_scalar<float>(a, 5));
    return AQC_OK;
}

// cfd/ideal_gas/symmetry/Mirror.cl:32-48 (sources are never mirrored particles themselves: no row is both
// read and written)
__global__ void __launch_bounds__(256)
k_ig_sym_set(const uint32_t* mirror_src, float* eint_in, float* deintdt_in, float* deintdt, uint32_t N)
{
    GID;
    const uint32_t s = mirror_src[i];
    if (s >= N)
        return;
    eint_in[i] = eint_in[s];
    deintdt[i] = deintdt_in[i] = deintdt_in[s];
}
int l_ig_sym_set(aqc_ctx* c, size_t, void* const* a)
{
    const uint32_t N = aqc_scalar<uint32_t>(a, 4);
    LAUNCH(c, k_ig_sym_set, N, (const uint32_t*)a[0], (float*)a[1], (float*)a[2], (float*)a[3], N);
    return AQC_OK;
}

#define IN(n, t) { n, t, AQC_ARG_ARRAY_IN }
#define OUT(n, t) { n, t, AQC_ARG_ARRAY_OUT }
#define RO(n, t) { n, t, AQC_ARG_ARRAY_RO }
#define SC(n, t) { n, t, AQC_ARG_SCALAR }

#define MPI_FIELDS_OUT                                                            \
    OUT("mpi_iset", "unsigned int*"), OUT("mpi_r", "vec*"), OUT("mpi_u", "vec*"),  \
    OUT("mpi_dudt", "vec*"), OUT("mpi_rho", "float*"), OUT("mpi_drhodt", "float*"), \
    OUT("mpi_m", "float*")
#define MPI_FIELDS_IN                                                             \
    IN("mpi_iset", "unsigned int*"), IN("mpi_r", "vec*"), IN("mpi_u", "vec*"),     \
    IN("mpi_dudt", "vec*"), IN("mpi_rho", "float*"), IN("mpi_drhodt", "float*"),   \
    IN("mpi_m", "float*")
aqc_registrar r_mpi_copy("cfd/MPI.cl", "copy", 0,
    { MPI_FIELDS_OUT, IN("iset", "unsigned int*"), IN("r", "vec*"), IN("u", "vec*"),
      IN("dudt", "vec*"), IN("rho", "float*"), IN("drhodt", "float*"), IN("m", "float*"),
      SC("N", "usize") }, l_mpi_copy);
aqc_registrar r_mpi_append("cfd/MPI.cl", "append", 0,
    { OUT("iset", "unsigned int*"), OUT("r", "vec*"), OUT("u", "vec*"), OUT("dudt", "vec*"),
      OUT("rho", "float*"), OUT("drhodt", "float*"), OUT("m", "float*"), OUT("imove", "int*"),
      IN("mpi_local_mask", "usize*"), MPI_FIELDS_IN, SC("mpi_rank", "unsigned int"),
      SC("nbuffer", "usize"), SC("N", "usize") }, l_mpi_append);
aqc_registrar r_mpi_remove("cfd/MPI.cl", "remove", 0,
    { OUT("imove", "int*"), OUT("r", "vec*"), OUT("u", "vec*"), OUT("dudt", "vec*"),
      OUT("m", "float*"), IN("mpi_local_mask", "usize*"), SC("mpi_rank", "unsigned int"),
      SC("domain_max", "vec"), SC("N", "usize") }, l_mpi_remove);
aqc_registrar r_mpi_backup_r("cfd/MPI.cl", "backup_r", 0,
    { IN("mpi_neigh_mask", "usize*"), IN("mpi_r", "vec*"), OUT("mpi_r_in", "vec*"),
      SC("r_max", "vec"), SC("mpi_rank", "unsigned int"), SC("N", "usize"),
      SC("n_radix", "usize") }, l_mpi_backup_r);
aqc_registrar r_mpi_sort("cfd/MPI.cl", "sort", 0,
    { IN("mpi_iset_in", "unsigned int*"), OUT("mpi_iset", "unsigned int*"), IN("mpi_r_in", "vec*"),
      OUT("mpi_r", "vec*"), IN("mpi_u_in", "vec*"), OUT("mpi_u", "vec*"),
      IN("mpi_rho_in", "float*"), OUT("mpi_rho", "float*"), IN("mpi_m_in", "float*"),
      OUT("mpi_m", "float*"), IN("mpi_id_sorted", "usize*"), SC("N", "usize") }, l_mpi_sort);
aqc_registrar r_mpi_eos("cfd/MPI.cl", "eos", 0,
    { OUT("mpi_iset", "unsigned int*"), OUT("mpi_rho", "float*"), OUT("mpi_p", "float*"),
      IN("refd", "float*"), SC("N", "usize"), SC("cs", "float"), SC("p0", "float") }, l_mpi_eos);
aqc_registrar r_mpi_copy_g("aqua/MPIdeltaSPH.cl", "copy_g", 0,
    { OUT("mpi_lap_p_corr", "vec*"), IN("lap_p_corr", "vec*"), SC("N", "usize") }, l_mpi_copy_g);
aqc_registrar r_mpi_sort_g("aqua/MPIdeltaSPH.cl", "sort_g", 0,
    { IN("mpi_lap_p_corr_in", "vec*"), OUT("mpi_lap_p_corr", "vec*"), IN("mpi_id_sorted", "usize*"),
      SC("N", "usize") }, l_mpi_sort_g);
aqc_registrar r_mpi_lmask("cfd/MPI/planes.cl", "local_mask", 0,
    { IN("imove", "int*"), IN("r", "vec*"), OUT("mpi_local_mask", "usize*"),
      SC("mpi_plane_r", "vec"), SC("mpi_plane_n", "vec"), SC("mpi_plane_proc", "unsigned int"),
      SC("N", "usize") }, l_mpi_local_mask);
aqc_registrar r_mpi_nmask("cfd/MPI/planes.cl", "neigh_mask", 0,
    { IN("imove", "int*"), IN("r", "vec*"), OUT("mpi_neigh_mask", "usize*"),
      SC("mpi_plane_r", "vec"), SC("mpi_plane_n", "vec"), SC("mpi_plane_proc", "unsigned int"),
      SC("N", "usize") }, l_mpi_neigh_mask);

aqc_registrar r_bi_gradp("cfd/Boundary/BI/GradP.cl", "freeslip", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("shepard", "float*"), IN("rho", "float*"),
      IN("lap_u", "vec*"), IN("dudt", "vec*"), OUT("grad_p", "vec*"), IN("visc_dyn", "float*"),
      IN("refd", "float*"), SC("N", "usize"), SC("g", "vec") }, l_bi_gradp);
aqc_registrar r_bi_ishep("cfd/Boundary/BI/InterpolationShepard.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("shepard", "float*"), OUT("rho", "float*"),
      OUT("p", "float*"), IN("refd", "float*"), SC("N", "usize"), SC("cs", "float"),
      SC("p0", "float") }, l_bi_interp_shepard);
aqc_registrar r_bi_shapply("cfd/Boundary/BI/Shepard.cl", "apply", 0,
    { IN("imove", "int*"), IN("shepard", "float*"), OUT("grad_p", "vec*"), OUT("lap_u", "vec*"),
      OUT("div_u", "float*"), SC("N", "usize"), SC("cs", "float") }, l_bi_shepard_apply);

aqc_registrar r_h_sensor("h_sensor.cl", "entry", 3,
    { IN("imove", "int*"), IN("r", "vec*"), OUT("h_sensorz", "float*"), SC("N", "uint"),
      SC("h_sensorx", "float"), SC("dr", "float") }, l_h_sensor);

#define STATE_COPY_ARGS                                                          \
    { IN("r", "vec*"), IN("u", "vec*"), IN("dudt", "vec*"), IN("rho", "float*"),   \
      IN("drhodt", "float*"), OUT("r_in", "vec*"), OUT("u_in", "vec*"),            \
      OUT("dudt_in", "vec*"), OUT("rho_in", "float*"), OUT("drhodt_in", "float*"), \
      SC("N", "usize") }
// (euler.cl:65-75 declares the state it only reads without const)
aqc_registrar r_eu_p("basic/time_scheme/euler.cl", "predictor", 0,
    { RO("r", "vec*"), RO("u", "vec*"), RO("dudt", "vec*"), RO("rho", "float*"), RO("drhodt", "float*"),
      OUT("r_in", "vec*"), OUT("u_in", "vec*"), OUT("dudt_in", "vec*"), OUT("rho_in", "float*"),
      OUT("drhodt_in", "float*"), SC("N", "usize") }, l_copy_state);
aqc_registrar r_mp_p("basic/time_scheme/midpoint.cl", "predictor", 0, STATE_COPY_ARGS, l_copy_state);
aqc_registrar r_eu_c("basic/time_scheme/euler.cl", "corrector", 0,
    { OUT("imove", "int*"), OUT("iset", "unsigned int*"), OUT("r", "vec*"), OUT("u", "vec*"),
      OUT("dudt", "vec*"), OUT("rho", "float*"), OUT("drhodt", "float*"), SC("N", "usize"),
      SC("dt", "float") }, l_euler_corrector);
aqc_registrar r_ie_p("basic/time_scheme/improved_euler.cl", "predictor", 0,
    { OUT("imove", "int*"), OUT("r", "vec*"), OUT("u", "vec*"), OUT("dudt", "vec*"),
      OUT("rho", "float*"), OUT("drhodt", "float*"), OUT("r_in", "vec*"), OUT("u_in", "vec*"),
      OUT("dudt_in", "vec*"), OUT("rho_in", "float*"), OUT("drhodt_in", "float*"),
      SC("N", "usize"), SC("dt", "float") }, l_ie_predictor);
aqc_registrar r_ie_c("basic/time_scheme/improved_euler.cl", "corrector", 0,
    { OUT("imove", "int*"), OUT("iset", "unsigned int*"), OUT("r", "vec*"), OUT("u", "vec*"),
      OUT("dudt", "vec*"), OUT("rho", "float*"), OUT("drhodt", "float*"), OUT("dudt_in", "vec*"),
      OUT("drhodt_in", "float*"), SC("N", "usize"), SC("dt", "float") }, l_ie_corrector);
aqc_registrar r_mp_m("basic/time_scheme/midpoint.cl", "midpoint", 0,
    { IN("imove", "int*"), IN("u_in", "vec*"), OUT("u", "vec*"), IN("dudt", "vec*"),
      IN("rho_in", "float*"), OUT("rho", "float*"), IN("drhodt", "float*"), SC("N", "usize"),
      SC("dt", "float") }, l_mp_midpoint);
aqc_registrar r_mp_mr("basic/time_scheme/midpoint.cl", "midpoint_r", 0,
    { IN("imove", "int*"), IN("r_in", "vec*"), OUT("r", "vec*"), IN("u", "vec*"), SC("N", "usize"),
      SC("dt", "float") }, l_mp_midpoint_r);
aqc_registrar r_mp_rx("basic/time_scheme/midpoint.cl", "relax", 0,
    { IN("imove", "int*"), OUT("dudt_in", "vec*"), OUT("dudt", "vec*"), OUT("drhodt_in", "float*"),
      OUT("drhodt", "float*"), SC("N", "usize"), SC("relax_midpoint", "float") }, l_mp_relax, 1ull << 6);
aqc_registrar r_mp_rs("basic/time_scheme/midpoint.cl", "residuals", 0,
    { IN("imove", "int*"), IN("m", "float*"), IN("u", "vec*"), IN("dudt_in", "vec*"),
      IN("dudt", "vec*"), IN("rho", "float*"), IN("p", "float*"), IN("drhodt_in", "float*"),
      IN("drhodt", "float*"), OUT("residual_midpoint", "float*"), SC("N", "usize") },
    l_mp_residuals);
aqc_registrar r_mp_c("basic/time_scheme/midpoint.cl", "corrector", 0,
    { IN("imove", "int*"), IN("r_in", "vec*"), OUT("r", "vec*"), IN("u_in", "vec*"),
      OUT("u", "vec*"), IN("dudt", "vec*"), IN("rho_in", "float*"), OUT("rho", "float*"),
      IN("drhodt", "float*"), SC("N", "usize"), SC("dt", "float") }, l_mp_corrector);
// (r of Velocity / Acceleration is declared without const and only read)
aqc_registrar r_mo_t("cfd/Motions/Transform.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), OUT("r", "vec*"), OUT("normal", "vec*"), OUT("tangent", "vec*"),
      SC("N", "usize"), SC("motion_iset", "unsigned int"), SC("motion_r", "vec"), SC("motion_a", "vec4") },
    l_motion_transform);
aqc_registrar r_mo_u("cfd/Motions/UnTransform.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), OUT("r", "vec*"), OUT("normal", "vec*"), OUT("tangent", "vec*"),
      SC("N", "usize"), SC("motion_iset", "unsigned int"), SC("motion_r_in", "vec"), SC("motion_a_in", "vec4") },
    l_motion_untransform);
aqc_registrar r_mo_v("cfd/Motions/Velocity.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), RO("r", "vec*"), OUT("u", "vec*"), SC("N", "usize"),
      SC("motion_iset", "unsigned int"), SC("motion_drdt", "vec"), SC("motion_a", "vec4"),
      SC("motion_dadt", "vec4") }, l_motion_velocity);
aqc_registrar r_mo_a("cfd/Motions/Acceleration.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), RO("r", "vec*"), OUT("dudt", "vec*"), SC("N", "usize"),
      SC("motion_iset", "unsigned int"), SC("motion_r", "vec"), SC("motion_ddrddt", "vec"),
      SC("motion_a", "vec4"), SC("motion_ddaddt", "vec4") }, l_motion_acceleration);
aqc_registrar r_en_p("cfd/Energy/Energy.cl", "power", 0,
    { OUT("energy_dekdt", "float*"), OUT("energy_depdt", "float*"), OUT("energy_decdt", "float*"),
      IN("imove", "int*"), IN("u", "vec*"), IN("rho", "float*"), IN("m", "float*"), IN("p", "float*"),
      IN("dudt", "vec*"), IN("drhodt", "float*"), SC("N", "usize"), SC("g", "vec") }, l_energy_power);
aqc_registrar r_en_e("cfd/Energy/Energy.cl", "energy", 0,
    { OUT("energy_ek", "float*"), OUT("energy_ep", "float*"), OUT("energy_ec", "float*"),
      IN("iset", "uint*"), IN("imove", "int*"), IN("r", "vec*"), IN("u", "vec*"), IN("rho", "float*"),
      IN("m", "float*"), IN("refd", "float*"), SC("N", "usize"), SC("g", "vec"), SC("cs", "float") },
    l_energy_energy);
// basic/time_scheme/adam_bashforth.cl: ::predictor declares the state it only reads without const, ::corrector
// everything; the history levels are read-only there
aqc_registrar r_ab_p("basic/time_scheme/adam_bashforth.cl", "predictor", 0,
    { RO("r", "vec*"), RO("u", "vec*"), RO("dudt", "vec*"), RO("rho", "float*"), RO("drhodt", "float*"),
      OUT("r_in", "vec*"), OUT("u_in", "vec*"), OUT("dudt_in", "vec*"), OUT("rho_in", "float*"),
      OUT("drhodt_in", "float*"), SC("N", "usize") }, l_copy_state);
aqc_registrar r_ab_s("basic/time_scheme/adam_bashforth.cl", "sort", 0,
    { IN("dudt_as1_in", "vec*"), OUT("dudt_as1", "vec*"), IN("dudt_as2_in", "vec*"), OUT("dudt_as2", "vec*"),
      IN("dudt_as3_in", "vec*"), OUT("dudt_as3", "vec*"), IN("dudt_as4_in", "vec*"), OUT("dudt_as4", "vec*"),
      IN("drhodt_as1_in", "float*"), OUT("drhodt_as1", "float*"), IN("drhodt_as2_in", "float*"),
      OUT("drhodt_as2", "float*"), IN("drhodt_as3_in", "float*"), OUT("drhodt_as3", "float*"),
      IN("drhodt_as4_in", "float*"), OUT("drhodt_as4", "float*"), IN("id_sorted", "usize*"), SC("N", "usize") },
    l_ab_sort);
aqc_registrar r_ab_c("basic/time_scheme/adam_bashforth.cl", "corrector", 0,
    { RO("imove", "int*"), RO("iset", "unsigned int*"), OUT("r", "vec*"), OUT("u", "vec*"), RO("dudt", "vec*"),
      OUT("rho", "float*"), RO("drhodt", "float*"), RO("dudt_as1", "vec*"), RO("drhodt_as1", "float*"),
      RO("dudt_as2", "vec*"), RO("drhodt_as2", "float*"), RO("dudt_as3", "vec*"), RO("drhodt_as3", "float*"),
      RO("dudt_as4", "vec*"), RO("drhodt_as4", "float*"), SC("N", "usize"), SC("dt", "float"),
      SC("iter", "unsigned int") }, l_ab_corrector);
aqc_registrar r_ab_pc("basic/time_scheme/adam_bashforth.cl", "postcorrector", 0,
    { IN("dudt_as1", "vec*"), IN("drhodt_as1", "float*"), IN("dudt_as2", "vec*"), IN("drhodt_as2", "float*"),
      IN("dudt_as3", "vec*"), IN("drhodt_as3", "float*"), IN("dudt", "vec*"), IN("drhodt", "float*"),
      OUT("dudt_as1_in", "vec*"), OUT("drhodt_as1_in", "float*"), OUT("dudt_as2_in", "vec*"),
      OUT("drhodt_as2_in", "float*"), OUT("dudt_as3_in", "vec*"), OUT("drhodt_as3_in", "float*"),
      OUT("dudt_as4_in", "vec*"), OUT("drhodt_as4_in", "float*"), SC("N", "usize") }, l_ab_postcorrector);
aqc_registrar r_en_k("cfd/Energy/EnergyKin.cl", "entry", 0,
    { OUT("energy_kin", "float*"), IN("imove", "int*"), IN("u", "vec*"), IN("m", "float*"), SC("N", "usize") },
    l_energy_kin);
// (imove, r, dudt, m of Forces.cl and id of IdInverse.cl are declared without const and only read)
aqc_registrar r_forces("cfd/Forces/Forces.cl", "entry", 0,
    { OUT("forces_f", "vec*"), OUT("forces_m", "vec4*"), RO("imove", "int*"), RO("r", "vec*"), RO("dudt", "vec*"),
      RO("m", "float*"), SC("N", "usize"), SC("g", "vec"), SC("forces_r", "vec") }, l_forces);
// cfd/Boundary/Symmetry/Mirror.cl (preset cfd/symmetry.xml)
aqc_registrar r_sym_drop("cfd/Boundary/Symmetry/Mirror.cl", "drop", 0,
    { OUT("imove", "int*"), OUT("r", "vec*"), SC("N", "usize"), SC("symmetry_r", "vec"), SC("symmetry_n", "vec"),
      SC("domain_max", "vec") }, l_sym_drop);
aqc_registrar r_sym_detect("cfd/Boundary/Symmetry/Mirror.cl", "detect", 0,
    { IN("imove", "int*"), IN("r_in", "vec*"), OUT("imirror", "uint*"), SC("N", "usize"), SC("symmetry_r", "vec"),
      SC("symmetry_n", "vec") }, l_sym_detect);
aqc_registrar r_sym_feed("cfd/Boundary/Symmetry/Mirror.cl", "feed", 0,
    { OUT("imove", "int*"), OUT("iset", "uint*"), IN("imirror", "uint*"), IN("imirror_invperm", "usize*"),
      OUT("mirror_src", "usize*"), OUT("normal", "vec*"), OUT("tangent", "vec*"), OUT("r_in", "vec*"),
      SC("N", "usize"), SC("nbuffer", "usize"), SC("symmetry_r", "vec"), SC("symmetry_n", "vec") }, l_sym_feed);
aqc_registrar r_sym_set("cfd/Boundary/Symmetry/Mirror.cl", "set", 0,
    { IN("mirror_src", "usize*"), OUT("m", "float*"), OUT("u_in", "vec*"), OUT("dudt_in", "vec*"),
      OUT("dudt", "vec*"), OUT("rho_in", "float*"), OUT("drhodt_in", "float*"), OUT("drhodt", "float*"),
      SC("N", "usize"), SC("symmetry_r", "vec"), SC("symmetry_n", "vec") }, l_sym_set);
aqc_registrar r_sym_sort("cfd/Boundary/Symmetry/Mirror.cl", "sort", 0,
    { IN("mirror_src_in", "usize*"), OUT("mirror_src", "usize*"), IN("id_sorted", "usize*"), SC("N", "usize") },
    l_sym_sort);
aqc_registrar r_rho_clamp("basic/DensityClamp.cl", "entry", 0,
    { OUT("rho_in", "float*"), SC("N", "usize"), SC("rho_min", "float"), SC("rho_max", "float") },
    l_density_clamp);
aqc_registrar r_id_inv("basic/IdInverse.cl", "entry", 0,
    { RO("id", "usize*"), OUT("id_inverse", "usize*"), SC("N", "usize") }, l_id_inverse);
aqc_registrar r_domain("basic/Domain.cl", "entry", 0,
    { OUT("imove", "int*"), OUT("r_in", "vec*"), OUT("u_in", "vec*"), OUT("dudt_in", "vec*"),
      OUT("m", "float*"), SC("N", "usize"), SC("domain_min", "vec"), SC("domain_max", "vec") },
    l_domain);
aqc_registrar r_sort1("basic/Sort.cl", "stage1", 0,
    { IN("id_in", "usize*"), OUT("id", "usize*"), IN("iset_in", "uint*"), OUT("iset", "uint*"),
      IN("imove_in", "int*"), OUT("imove", "int*"), IN("r_in", "vec*"), OUT("r", "vec*"),
      IN("normal_in", "vec*"), OUT("normal", "vec*"), IN("tangent_in", "vec*"),
      OUT("tangent", "vec*"), IN("id_sorted", "usize*"), SC("N", "usize") }, l_sort_stage1);
aqc_registrar r_sort2("basic/Sort.cl", "stage2", 0,
    { IN("rho_in", "float*"), OUT("rho", "float*"), IN("m_in", "float*"), OUT("m", "float*"),
      IN("u_in", "vec*"), OUT("u", "vec*"), IN("dudt", "vec*"), OUT("dudt_in", "vec*"),
      IN("drhodt", "float*"), OUT("drhodt_in", "float*"), IN("id_sorted", "usize*"),
      SC("N", "usize") }, l_sort_stage2);
aqc_registrar r_eos("basic/EOS.cl", "entry", 0,
    { RO("iset", "unsigned int*"), RO("imove", "int*"), RO("rho", "float*"), OUT("p", "float*"),
      IN("refd", "float*"), SC("N", "usize"), SC("cs", "float"), SC("p0", "float") }, l_eos);
aqc_registrar r_binormal("basic/Binormal.cl", "entry", 0,
    { IN("normal", "vec*"), OUT("tangent", "vec*"), OUT("binormal", "vec*"), SC("N", "usize") },
    l_binormal);
aqc_registrar r_rates("cfd/Rates.cl", "entry", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), IN("rho", "float*"), IN("grad_p", "vec*"),
      IN("lap_u", "vec*"), IN("div_u", "float*"), OUT("dudt", "vec*"), OUT("drhodt", "float*"),
      IN("visc_dyn", "float*"), SC("N", "usize"), SC("g", "vec") }, l_rates);
aqc_registrar r_timestep("cfd/TimeStep.cl", "entry", 0,
    { IN("imove", "int*"), IN("u", "vec*"), OUT("dt_var", "float*"), SC("N", "usize"),
      SC("dt", "float"), SC("dt_min", "float"), SC("courant", "float"), SC("dt_Ma", "float"),
      SC("h", "float") }, l_timestep);
aqc_registrar r_sens_rn("cfd/SensorsRenormalization.cl", "entry", 0,
    { IN("imove", "int*"), IN("shepard", "float*"), OUT("u", "vec*"), OUT("rho", "float*"),
      OUT("p", "float*"), SC("N", "usize"), SC("dt", "float"), SC("g", "vec") }, l_sensors_renorm);
aqc_registrar r_ds_simple("cfd/deltaSPH.cl", "simple", 0,
    { IN("iset", "unsigned int*"), IN("imove", "int*"), OUT("lap_p_corr", "vec*"),
      IN("refd", "float*"), SC("N", "usize"), SC("g", "vec") }, l_dsph_simple);
aqc_registrar r_ds_fmls("cfd/deltaSPH.cl", "full_mls", 0,
    { IN("imove", "int*"), IN("mls", "matrix*"), OUT("lap_p_corr", "vec*"), SC("N", "usize") },
    l_dsph_full_mls);
aqc_registrar r_ds_apply("cfd/deltaSPH.cl", "deltaSPH", 0,
    { IN("iset", "unsigned int*"), IN("imove", "int*"), IN("rho", "float*"), IN("lap_p", "float*"),
      OUT("drhodt", "float*"), IN("refd", "float*"), IN("delta", "float*"), SC("N", "usize"),
      SC("dt", "float") }, l_dsph_apply);
aqc_registrar r_mls_inv("basic/MLS.cl", "mls_inv", 0,
    { IN("imove", "int*"), OUT("mls", "matrix*"), SC("N", "usize"), SC("mls_imove", "uint") },
    l_mls_inv);
aqc_registrar r_bie_r("cfd/Boundary/BIe/Rates.cl", "entry", 0,
    { IN("imove", "int*"), IN("rho", "float*"), IN("p", "float*"), IN("u", "vec*"),
      IN("grad_w_bi", "vec*"), IN("div_u_bi", "float*"), OUT("grad_p", "vec*"),
      OUT("div_u", "float*"), SC("N", "usize") }, l_bie_rates);
aqc_registrar r_bie_fp("cfd/Boundary/BIe/Rates.cl", "filter_press", 0,
    { IN("iset", "uint*"), IN("imove", "int*"), OUT("p", "float*"),
      SC("forces_iset", "unsigned int"), SC("N", "usize") }, l_bie_filter_press);
aqc_registrar r_bie_frp("cfd/Boundary/BIe/Rates.cl", "force_press", 0,
    { IN("imove", "int*"), IN("r", "vec*"), IN("normal", "vec*"), IN("m", "float*"),
      IN("p", "float*"), OUT("force_p", "vec*"), OUT("moment_p", "vec4*"), SC("forces_r", "vec"),
      SC("N", "usize") }, l_bie_force_press);
aqc_registrar r_bie_fb("cfd/Boundary/BIe/ElasticBounce.cl", "force_bound", 0,
    { IN("imove", "int*"), IN("m", "float*"), IN("dudt_preelastic", "vec*"),
      IN("dudt_elastic", "vec*"), OUT("force_elastic", "vec*"), SC("N", "usize") },
    l_bie_force_bound);
aqc_registrar r_sb_count("basic/SetBuffer.cl", "count", 0,
    { IN("imove", "int*"), OUT("ibuffer", "unsigned int*"), SC("N", "usize") }, l_setbuffer_count);
aqc_registrar r_sb_set("basic/SetBuffer.cl", "set_imove", 0,
    { OUT("imove", "int*"), SC("N", "usize") }, l_setbuffer_set_imove);

// cfd/Boundary/Inlet/Inlet.cl, Outlet/Outlet.cl, Portal/Mirror.cl (presets cfd/inlet.xml, outlet.xml, portal.xml)
aqc_registrar r_inlet_feed("cfd/Boundary/Inlet/Inlet.cl", "feed", 0,
    { OUT("imove", "int*"), RO("iset", "unsigned int*"), OUT("r", "vec*"), OUT("u", "vec*"), OUT("dudt", "vec*"),
      OUT("rho", "float*"), OUT("drhodt", "float*"), OUT("m", "float*"), OUT("p", "float*"), IN("refd", "float*"),
      SC("N", "usize"), SC("nbuffer", "usize"), SC("dt", "float"), SC("cs", "float"), SC("p0", "float"),
      SC("g", "vec"), SC("dr", "float"), SC("inlet_r", "vec"), SC("inlet_ru", "vec"), SC("inlet_rv", "vec"),
      SC("inlet_N", "svec2"), SC("inlet_n", "vec"), SC("inlet_U", "float"), SC("inlet_rFS", "vec"),
      SC("inlet_R", "float"), SC("inlet_starving", "int") }, l_inlet_feed);
aqc_registrar r_inlet_rates("cfd/Boundary/Inlet/Inlet.cl", "rates", 0,
    { RO("imove", "int*"), RO("r", "vec*"), OUT("u", "vec*"), OUT("dudt", "vec*"), OUT("drhodt", "float*"),
      SC("N", "usize"), SC("inlet_r", "vec"), SC("inlet_U", "float"), SC("inlet_n", "vec") }, l_inlet_rates);
aqc_registrar r_outlet_rates("cfd/Boundary/Outlet/Outlet.cl", "rates", 0,
    { RO("imove", "int*"), RO("iset", "unsigned int*"), RO("r", "vec*"), OUT("u", "vec*"), OUT("rho", "float*"),
      OUT("p", "float*"), OUT("dudt", "vec*"), OUT("dudt_in", "vec*"), OUT("drhodt", "float*"),
      OUT("drhodt_in", "float*"), IN("refd", "float*"), SC("N", "usize"), SC("cs", "float"), SC("p0", "float"),
      SC("g", "vec"), SC("outlet_r", "vec"), SC("outlet_n", "vec"), SC("outlet_U", "float"),
      SC("outlet_rFS", "vec") }, l_outlet_rates);
aqc_registrar r_outlet_feed("cfd/Boundary/Outlet/Outlet.cl", "feed", 0,
    { OUT("imove", "int*"), OUT("r_in", "vec*"), SC("N", "usize"), SC("domain_max", "vec"), SC("outlet_r", "vec"),
      SC("outlet_n", "vec") }, l_outlet_feed);
aqc_registrar r_portal_mirror("cfd/Boundary/Portal/Mirror.cl", "mirror", 0,
    { OUT("r", "vec*"), OUT("imirrored", "int*"), OUT("icell", "usize*"), SC("N", "usize"), SC("portal_in_r", "vec"),
      SC("portal_out_r", "vec"), SC("portal_n", "vec"), SC("r_min", "vec"), SC("n_cells", "uivec4") },
    l_portal_mirror);
aqc_registrar r_portal_unmirror("cfd/Boundary/Portal/Mirror.cl", "unmirror", 0,
    { OUT("r", "vec*"), IN("imirrored", "int*"), SC("N", "usize"), SC("portal_in_r", "vec"),
      SC("portal_out_r", "vec"), SC("portal_n", "vec") }, l_portal_unmirror);
aqc_registrar r_portal_teleport("cfd/Boundary/Portal/Mirror.cl", "teleport", 0,
    { OUT("r", "vec*"), SC("N", "usize"), SC("portal_in_r", "vec"), SC("portal_out_r", "vec"),
      SC("portal_n", "vec") }, l_portal_teleport);

aqc_registrar r_ig_eos("cfd/ideal_gas/EOS.cl", "entry", 0,
    { IN("iset", "unsigned int*"), IN("imove", "int*"), IN("rho", "float*"), IN("eint", "float*"),
      OUT("p", "float*"), IN("gamma", "float*"), SC("N", "usize") }, l_ig_eos);
aqc_registrar r_ig_rates("cfd/ideal_gas/Rates.cl", "entry", 0,
    { IN("iset", "unsigned int*"), IN("imove", "int*"), IN("rho", "float*"), IN("p", "float*"),
      IN("div_u", "float*"), OUT("deintdt", "float*"), SC("N", "usize") }, l_ig_rates);
aqc_registrar r_ig_sort("cfd/ideal_gas/Sort.cl", "entry", 0,
    { IN("eint_in", "float*"), OUT("eint", "float*"), IN("deintdt", "float*"), OUT("deintdt_in", "float*"),
      IN("id_sorted", "usize*"), SC("N", "usize") }, l_ig_sort);
aqc_registrar r_ig_dt("cfd/ideal_gas/TimeStep.cl", "entry", 0,
    { OUT("dt_var", "float*"), IN("imove", "int*"), IN("iset", "unsigned int*"), IN("u", "vec*"),
      IN("dudt", "vec*"), IN("rho", "float*"), IN("p", "float*"), IN("m", "float*"), SC("N", "usize"),
      SC("dt", "float"), SC("dt_min", "float"), SC("courant", "float"), SC("h", "float"),
      IN("div_u", "float*"), IN("grad_p", "vec*"), IN("gamma", "float*") }, l_ig_timestep);
aqc_registrar r_ig_rrates("cfd/ideal_gas/riemann/Rates.cl", "entry", 0,
    { IN("iset", "unsigned int*"), IN("imove", "int*"), IN("work_density", "float*"), OUT("deintdt", "float*"),
      SC("N", "usize") }, l_ig_riemann_rates);
aqc_registrar r_ig_eu_p("cfd/ideal_gas/time_scheme/euler.cl", "predictor", 0,
    { IN("eint", "float*"), IN("deintdt", "float*"), OUT("eint_in", "float*"), OUT("deintdt_in", "float*"),
      SC("N", "usize") }, l_ig_mp_predictor);
aqc_registrar r_ig_eu_c("cfd/ideal_gas/time_scheme/euler.cl", "corrector", 0,
    { IN("imove", "int*"), OUT("eint", "float*"), IN("deintdt", "float*"), SC("N", "uint"), SC("dt", "float") },
    l_ig_euler_corrector);
aqc_registrar r_ig_ie_p("cfd/ideal_gas/time_scheme/improved_euler.cl", "predictor", 0,
    { RO("imove", "int*"), IN("eint", "float*"), IN("deintdt", "float*"), OUT("eint_in", "float*"),
      OUT("deintdt_in", "float*"), SC("N", "usize"), SC("dt", "float") }, l_ig_ie_predictor);
aqc_registrar r_ig_ie_c("cfd/ideal_gas/time_scheme/improved_euler.cl", "corrector", 0,
    { IN("imove", "int*"), IN("deintdt", "float*"), IN("deintdt_in", "float*"), OUT("eint", "float*"),
      SC("N", "usize"), SC("dt", "float") }, l_ig_ie_corrector);
aqc_registrar r_ig_sym("cfd/ideal_gas/symmetry/Mirror.cl", "set", 0,
    { IN("mirror_src", "usize*"), OUT("eint_in", "float*"), OUT("deintdt_in", "float*"), OUT("deintdt", "float*"),
      SC("N", "usize") }, l_ig_sym_set);
aqc_registrar r_ig_mp_p("cfd/ideal_gas/time_scheme/midpoint.cl", "predictor", 0,
    { IN("eint", "float*"), IN("deintdt", "float*"), OUT("eint_in", "float*"), OUT("deintdt_in", "float*"),
      SC("N", "usize") }, l_ig_mp_predictor);
aqc_registrar r_ig_mp_m("cfd/ideal_gas/time_scheme/midpoint.cl", "midpoint", 0,
    { IN("imove", "int*"), IN("eint_in", "float*"), IN("deintdt", "float*"), OUT("eint", "float*"),
      SC("N", "usize"), SC("dt", "float") }, l_ig_mp_midpoint);
aqc_registrar r_ig_mp_c("cfd/ideal_gas/time_scheme/midpoint.cl", "corrector", 0,
    { IN("imove", "int*"), IN("eint_in", "float*"), IN("deintdt", "float*"), OUT("eint", "float*"),
      SC("N", "usize"), SC("dt", "float") }, l_ig_mp_corrector);
aqc_registrar r_ig_mp_r("cfd/ideal_gas/time_scheme/midpoint.cl", "relax", 0,
    { IN("imove", "int*"), IN("deintdt_in", "float*"), OUT("deintdt", "float*"), SC("N", "usize"),
      SC("relax_midpoint", "float") }, l_ig_mp_relax, 1ull << 4);

} // namespace
