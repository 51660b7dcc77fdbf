/* aqo_kernels.c -- CPU ORACLE (test infrastructure, not product code).
 * Restates the preset script kernels of AQUAgpusph 5.0.4 that sit on the
 * per-time-step hot path (paths below are under resources/Scripts/).
 * See aqo.h for conventions.  Summation order follows the reference exactly:
 * cells x-outer / y / z-inner, ascending sorted index inside a cell
 * (types/3D.h:197-219, types/2D.h:174-193). */
#include "aqo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Outer particle loops honour a per-thread [lo, hi) window (aqo_set_range) so a
 * caller can spread one kernel over several host threads; default = everything.
 * Every kernel below only writes row i inside its i-loop, so windows are
 * independent. */
static __thread aqo_usize aqo_lo = 0, aqo_hi = 0xFFFFFFFFu;
void aqo_set_range(aqo_usize lo, aqo_usize hi)
{
    aqo_lo = lo;
    aqo_hi = hi;
}
#define AQO_FOR_I(n)                                                           \
    for (aqo_usize i = aqo_lo, i_end__ = ((n) < aqo_hi ? (n) : aqo_hi); i < i_end__; i++)

#define VS(dims) ((dims) == 3 ? 4 : 2)
#define MS(dims) ((dims) == 3 ? 16 : 4)
#define iM_PI 0.318309886f /* KernelFunctions/Wendland3D.hcl:34-39 */

/* BEGIN_NEIGHS / END_NEIGHS, types/3D.h:197-219 and types/2D.h:174-193.
 * Cell ids use unsigned wrap-around arithmetic like the OpenCL source. */
#define NEIGHS_BEGIN(L, i, dims)                                               \
    {                                                                          \
        const aqo_usize c_i__ = (L)->icell[i];                                 \
        const aqo_usize nx__ = (L)->ncells[0], ny__ = (L)->ncells[1];          \
        const int kz__ = ((dims) == 3) ? 1 : 0;                                \
        for (int ci__ = -1; ci__ <= 1; ci__++)                                 \
            for (int cj__ = -1; cj__ <= 1; cj__++)                             \
                for (int ck__ = -kz__; ck__ <= kz__; ck__++) {                 \
                    const aqo_usize c_j__ =                                    \
                        c_i__ + (aqo_usize)ci__ + (aqo_usize)cj__ * nx__ +     \
                        (aqo_usize)ck__ * nx__ * ny__;                         \
                    for (aqo_usize j = (L)->ihoc[c_j__];                       \
                         (j < (L)->N) && ((L)->icell[j] == c_j__); j++) {
#define NEIGHS_END                                                             \
                    }                                                          \
                }                                                              \
    }

static inline float dotv(const float* a, const float* b, int dims)
{
    float s = a[0] * b[0] + a[1] * b[1];
    if (dims == 3)
        s += a[2] * b[2];
    return s;
}

/* KernelFunctions/Wendland3D.hcl:44-66, Wendland2D.hcl:44-66 */
static inline float kernelW(float q, int dims)
{
    const float wcon = (dims == 3 ? 0.08203125f : 0.109375f) * iM_PI;
    return wcon * (1.f + 2.f * q) * (2.f - q) * (2.f - q) * (2.f - q) * (2.f - q);
}
static inline float kernelF(float q, int dims)
{
    const float wcon = (dims == 3 ? 0.8203125f : 1.09375f) * iM_PI;
    return wcon * (2.f - q) * (2.f - q) * (2.f - q);
}

/* r_ij = r_j - r_i, q = length(r_ij) / H; returns 0 when q >= SUPPORT */
static inline int pair_q(const aqo_defs* D, const float* ri, const float* rj,
                         float* r_ij, float* q)
{
    for (int d = 0; d < D->dims; d++)
        r_ij[d] = rj[d] - ri[d];
    *q = sqrtf(dotv(r_ij, r_ij, D->dims)) / D->H;
    return *q < D->SUPPORT;
}

/* ======================= element-wise kernels ============================ */

/* basic/EOS.cl:57-73; EXCLUDED_PARTICLE = (imove <= 0) && (imove != -1) */
void aqo_eos(const aqo_usize* iset, const int* imove, const float* rho,
             float* p, const float* refd, aqo_usize N, float cs, float p0)
{
    AQO_FOR_I(N) {
        if ((imove[i] <= 0) && (imove[i] != -1))
            continue;
        p[i] = p0 + cs * cs * (rho[i] - refd[iset[i]]);
    }
}

/* cfd/Rates.cl:55-77 (whole vec, incl. w in 3D) */
void aqo_rates(const aqo_usize* iset, const int* imove, const float* rho,
               const float* grad_p, const float* lap_u, const float* div_u,
               float* dudt, float* drhodt, const float* visc_dyn, aqo_usize N,
               const float* g, int dims)
{
    (void)rho;
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        const float mu = visc_dyn[iset[i]];
        for (int c = 0; c < vs; c++)
            dudt[(size_t)i * vs + c] = -grad_p[(size_t)i * vs + c] +
                                       mu * lap_u[(size_t)i * vs + c] + g[c];
        drhodt[i] = -div_u[i];
    }
}

/* cfd/TimeStep.cl:56-77; length(u[i]) is over the whole vec (4 comps in 3D) */
void aqo_timestep(const int* imove, const float* u, float* dt_var, aqo_usize N,
                  float dt, float dt_min, float courant, float dt_Ma, float h,
                  int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0) {
            dt_var[i] = dt;
            continue;
        }
        const float* ui = u + (size_t)i * vs;
        float s = ui[0] * ui[0] + ui[1] * ui[1];
        if (dims == 3)
            s = s + ui[2] * ui[2] + ui[3] * ui[3];
        const float dr_max = dt_Ma * h;
        const float dt_u = courant * dr_max / sqrtf(s);
        dt_var[i] = fmaxf(fminf(dt, dt_u), dt_min);
    }
}

/* cfd/variableTimeStep.xml:11-16: reduction "c = min(a, b)" null INFINITY */
float aqo_reduce_min(const float* v, aqo_usize N)
{
    float m = INFINITY;
    AQO_FOR_I(N)
        m = fminf(m, v[i]);
    return m;
}
float aqo_reduce_max(const float* v, aqo_usize N)
{
    float m = -INFINITY;
    AQO_FOR_I(N)
        m = fmaxf(m, v[i]);
    return m;
}
aqo_usize aqo_reduce_max_u32(const aqo_usize* v, aqo_usize N)
{
    aqo_usize m = 0;
    AQO_FOR_I(N)
        m = v[i] > m ? v[i] : m;
    return m;
}

/* Reduction.cl.in:35-66 + Reduction.cpp:376-436: each pass reduces groups of
 * wg consecutive items with a halving tree (lmem[t] += lmem[t + s],
 * s = wg/2 ... 1), out-of-range items hold the identity; passes repeat on the
 * per-group results until one value is left. */
void aqo_reduce_sum_vec_tree(const float* v, aqo_usize N, int ncomp,
                             aqo_usize wg, float* out)
{
    size_t n = N;
    float* cur = (float*)malloc(sizeof(float) * ncomp * (n ? n : 1));
    memcpy(cur, v, sizeof(float) * ncomp * n);
    float* lmem = (float*)malloc(sizeof(float) * ncomp * wg);
    if (n == 0)
        for (int c = 0; c < ncomp; c++)
            cur[c] = 0.f;
    while (n > 1) {
        const size_t groups = (n + wg - 1) / wg;
        for (size_t g = 0; g < groups; g++) {
            for (aqo_usize t = 0; t < wg; t++)
                for (int c = 0; c < ncomp; c++) {
                    const size_t gid = g * wg + t;
                    lmem[t * ncomp + c] = gid < n ? cur[gid * ncomp + c] : 0.f;
                }
            for (aqo_usize s = wg / 2; s > 0; s >>= 1)
                for (aqo_usize t = 0; t < s; t++)
                    for (int c = 0; c < ncomp; c++)
                        lmem[t * ncomp + c] =
                            lmem[t * ncomp + c] + lmem[(t + s) * ncomp + c];
            for (int c = 0; c < ncomp; c++)
                cur[g * ncomp + c] = lmem[c];
        }
        n = groups;
    }
    for (int c = 0; c < ncomp; c++)
        out[c] = cur[c];
    free(cur);
    free(lmem);
}
float aqo_reduce_sum_tree(const float* v, aqo_usize N, aqo_usize wg)
{
    float out;
    aqo_reduce_sum_vec_tree(v, N, 1, wg, &out);
    return out;
}

/* basic/Domain.cl:48-90 */
void aqo_domain(int* imove, float* r_in, float* u_in, float* dudt_in, float* m,
                aqo_usize N, const float* domain_min, const float* domain_max,
                int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= -255)
            continue;
        const float* c = r_in + (size_t)i * vs;
        int out = 0;
        for (int d = 0; d < dims; d++)
            out |= isnan(c[d]) || isinf(c[d]) || (c[d] < domain_min[d]) ||
                   (c[d] > domain_max[d]);
        if (!out)
            continue;
        imove[i] = -256;
        m[i] = 0.f;
        for (int d = 0; d < vs; d++) {
            u_in[(size_t)i * vs + d] = 0.f;
            dudt_in[(size_t)i * vs + d] = 0.f;
            r_in[(size_t)i * vs + d] = domain_max[d];
        }
    }
}

/* basic/Binormal.cl:37-53 */
void aqo_binormal(const float* normal, float* tangent, float* binormal,
                  aqo_usize N, int dims)
{
    AQO_FOR_I(N) {
        if (dims == 2) {
            binormal[2 * (size_t)i] = 0.f;
            binormal[2 * (size_t)i + 1] = 0.f;
            tangent[2 * (size_t)i] = normal[2 * (size_t)i + 1];
            tangent[2 * (size_t)i + 1] = -normal[2 * (size_t)i];
        } else {
            const float* a = normal + 4 * (size_t)i;
            const float* b = tangent + 4 * (size_t)i;
            float* o = binormal + 4 * (size_t)i;
            /* OpenCL cross(float4, float4): w = 0 */
            const float x = a[1] * b[2] - a[2] * b[1];
            const float y = a[2] * b[0] - a[0] * b[2];
            const float z = a[0] * b[1] - a[1] * b[0];
            o[0] = x;
            o[1] = y;
            o[2] = z;
            o[3] = 0.f;
        }
    }
}

/* ----------------------------- time schemes ------------------------------ */
static void copy_state(const float* r, const float* u, const float* dudt,
                       const float* rho, const float* drhodt, float* r_in,
                       float* u_in, float* dudt_in, float* rho_in,
                       float* drhodt_in, aqo_usize N, int dims)
{
    const size_t vb = sizeof(float) * VS(dims) * (size_t)N;
    memcpy(dudt_in, dudt, vb);
    memcpy(u_in, u, vb);
    memcpy(r_in, r, vb);
    memcpy(drhodt_in, drhodt, sizeof(float) * (size_t)N);
    memcpy(rho_in, rho, sizeof(float) * (size_t)N);
}

/* basic/time_scheme/euler.cl:65-87 */
void aqo_euler_predictor(const float* r, const float* u, const float* dudt,
                         const float* rho, const float* drhodt, float* r_in,
                         float* u_in, float* dudt_in, float* rho_in,
                         float* drhodt_in, aqo_usize N, int dims)
{
    copy_state(r, u, dudt, rho, drhodt, r_in, u_in, dudt_in, rho_in, drhodt_in,
               N, dims);
}

/* basic/time_scheme/euler.cl:105-124 */
void aqo_euler_corrector(const int* imove, float* r, float* u,
                         const float* dudt, float* rho, const float* drhodt,
                         aqo_usize N, float dt, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            r[k] += dt * u[k] + 0.5f * dt * dt * dudt[k];
            u[k] += dt * dudt[k];
        }
        rho[i] += dt * drhodt[i];
    }
}

/* basic/time_scheme/improved_euler.cl:75-103 */
void aqo_ie_predictor(const int* imove, const float* r, const float* u,
                      const float* dudt, const float* rho, const float* drhodt,
                      float* r_in, float* u_in, float* dudt_in, float* rho_in,
                      float* drhodt_in, aqo_usize N, float dt, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        const float DT = (imove[i] <= 0) ? 0.f : dt;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            dudt_in[k] = dudt[k];
            u_in[k] = u[k] + DT * dudt[k];
            r_in[k] = r[k] + DT * u[k] + 0.5f * DT * DT * dudt[k];
        }
        drhodt_in[i] = drhodt[i];
        rho_in[i] = rho[i] + DT * drhodt[i];
    }
}

/* basic/time_scheme/improved_euler.cl:125-147 */
void aqo_ie_corrector(const int* imove, float* r, float* u, const float* dudt,
                      float* rho, const float* drhodt, const float* dudt_in,
                      const float* drhodt_in, aqo_usize N, float dt, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        const float DT = 0.5f * dt;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            u[k] += DT * (dudt[k] - dudt_in[k]);
            r[k] += DT * DT * (dudt[k] - dudt_in[k]);
        }
        rho[i] += DT * (drhodt[i] - drhodt_in[i]);
    }
}

/* basic/time_scheme/midpoint.cl:53-75 */
void aqo_mp_predictor(const float* r, const float* u, const float* dudt,
                      const float* rho, const float* drhodt, float* r_in,
                      float* u_in, float* dudt_in, float* rho_in,
                      float* drhodt_in, aqo_usize N, int dims)
{
    copy_state(r, u, dudt, rho, drhodt, r_in, u_in, dudt_in, rho_in, drhodt_in,
               N, dims);
}

/* basic/time_scheme/midpoint.cl:93-111 */
void aqo_mp_midpoint(const int* imove, const float* u_in, float* u,
                     const float* dudt, const float* rho_in, float* rho,
                     const float* drhodt, aqo_usize N, float dt, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            u[k] = u_in[k] + 0.5f * dt * dudt[k];
        }
        rho[i] = rho_in[i] + 0.5f * dt * drhodt[i];
    }
}

/* basic/time_scheme/midpoint.cl:141-157 */
void aqo_mp_relax(const int* imove, const float* dudt_in, float* dudt,
                  const float* drhodt_in, float* drhodt, aqo_usize N,
                  float relax, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            dudt[k] = relax * dudt_in[k] + (1.f - relax) * dudt[k];
        }
        drhodt[i] = relax * drhodt_in[i] + (1.f - relax) * drhodt[i];
    }
}

/* basic/time_scheme/midpoint.cl:159-184; dot() over the whole vec */
void aqo_mp_residuals(const int* imove, const float* m, const float* u,
                      const float* dudt_in, const float* dudt, const float* rho,
                      const float* p, const float* drhodt_in,
                      const float* drhodt, float* residual, aqo_usize N,
                      int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0) {
            residual[i] = 0.f;
            continue;
        }
        const float rho2 = rho[i] * rho[i];
        float d = 0.f;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            const float t = u[k] * (dudt[k] - dudt_in[k]);
            d = (c == 0) ? t : d + t;
        }
        residual[i] = m[i] * (fabsf(d) +
                              fabsf(p[i] / rho2 * (drhodt[i] - drhodt_in[i])));
    }
}

/* basic/time_scheme/midpoint.cl:206-227 */
void aqo_mp_corrector(const int* imove, const float* r_in, float* r,
                      const float* u_in, float* u, const float* dudt,
                      const float* rho_in, float* rho, const float* drhodt,
                      aqo_usize N, float dt, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            r[k] = r_in[k] + dt * u_in[k] + 0.5f * dt * dt * dudt[k];
            u[k] = u_in[k] + dt * dudt[k];
        }
        rho[i] = rho_in[i] + dt * drhodt[i];
    }
}

/* ========================== neighbour sweeps ============================= */

/* cfd/Interactions.cl:60-145, __LAP_FORMULATION__ == __LAP_MONAGHAN__
 * (cfd.xml:52-54), __CLEARY__ = 8 (2D) / 10 (3D) (:33-39).  Built with
 * LOCAL_MEM_SIZE, i.e. the outputs are overwritten, not accumulated
 * (:140-144); only the XYZ components are written (w untouched). */
static void interactions_impl(const aqo_defs* D, const aqo_ll* L, const int* imove,
                              const float* r, const float* u, const float* rho,
                              const float* m, const float* p, float* grad_p,
                              float* lap_u, float* div_u, int morris)
{
    const int dims = D->dims, vs = VS(dims);
    const float cleary = (dims == 3) ? 10.f : 8.f;
    const float H = D->H;
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* u_i = u + (size_t)i * vs;
        const float p_i = p[i], rho_i = rho[i];
        float gp[3] = { 0.f, 0.f, 0.f }, lu[3] = { 0.f, 0.f, 0.f }, du = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (i == j)
                continue;
            if (imove[j] != 1)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float rho_j = rho[j], p_j = p[j];
            float u_ij[3];
            for (int d = 0; d < dims; d++)
                u_ij[d] = u[(size_t)j * vs + d] - u_i[d];
            const float udr = dotv(u_ij, r_ij, dims);
            const float f_ij = kernelF(q, dims) * D->CONF * m[j];
            const float pf = (p_i + p_j) / (rho_i * rho_j) * f_ij;
            if (morris) { /* :130-131 */
                const float lf = f_ij * 2.f / (rho_i * rho_j);
                for (int d = 0; d < dims; d++) {
                    gp[d] += pf * r_ij[d];
                    lu[d] += lf * u_ij[d];
                }
            } else { /* :127-129 */
                const float r2 = (q * q + 0.01f) * H * H;
                const float lf = f_ij * cleary * udr / (r2 * rho_i * rho_j);
                for (int d = 0; d < dims; d++) {
                    gp[d] += pf * r_ij[d];
                    lu[d] += lf * r_ij[d];
                }
            }
            du += udr * f_ij * rho_i / rho_j;
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++) {
            grad_p[(size_t)i * vs + d] = gp[d];
            lap_u[(size_t)i * vs + d] = lu[d];
        }
        div_u[i] = du;
    }
}

void aqo_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove,
                      const float* r, const float* u, const float* rho,
                      const float* m, const float* p, float* grad_p,
                      float* lap_u, float* div_u)
{
    interactions_impl(D, L, imove, r, u, rho, m, p, grad_p, lap_u, div_u, 0);
}

/* cfd/Interactions.cl:60-145 with __LAP_FORMULATION__ == __LAP_MORRIS__ (:130-131; the <Define> of
 * examples/2D/taylor_green and cylinder_inside_channel) */
void aqo_interactions_morris(const aqo_defs* D, const aqo_ll* L, const int* imove,
                             const float* r, const float* u, const float* rho,
                             const float* m, const float* p, float* grad_p,
                             float* lap_u, float* div_u)
{
    interactions_impl(D, L, imove, r, u, rho, m, p, grad_p, lap_u, div_u, 1);
}

/* basic/Shepard.cl:76-125 (cfd_mode 0, EXCLUDED = imove >= 3) and
 * cfd/Shepard.cl:29-35 (cfd_mode 1, EXCLUDED = imove != 1).  The self term is
 * included (no i == j test). */
void aqo_shepard(const aqo_defs* D, const aqo_ll* L, int cfd_mode,
                 const int* imove, const float* r, const float* rho,
                 const float* m, float* shepard)
{
    const int dims = D->dims, vs = VS(dims);
#define SH_EXCL(k) (cfd_mode ? (imove[k] != 1) : (imove[k] >= 3))
    AQO_FOR_I(L->N) {
        if ((imove[i] < -3) || ((imove[i] > 0) && SH_EXCL(i)))
            continue;
        const float* r_i = r + (size_t)i * vs;
        float s = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (SH_EXCL(j))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            s += kernelW(q, dims) * D->CONW * m[j] / rho[j];
        }
        NEIGHS_END
        shepard[i] = s;
    }
#undef SH_EXCL
}

/* basic/neighs.cl:52-91 */
void aqo_neighs(const aqo_ll* L, const int* imove, aqo_usize* n_neighs,
                aqo_usize neighs_limit, int dims)
{
    AQO_FOR_I(L->N) {
        if (imove[i] <= -255) {
            n_neighs[i] = 0;
            continue;
        }
        aqo_usize n = 0;
        NEIGHS_BEGIN(L, i, dims)
        {
            n += 1;
            if (n >= neighs_limit)
                goto done;
        }
        NEIGHS_END
    done:
        n_neighs[i] = n;
    }
}

/* cfd/Sensors.cl:57-130 */
void aqo_sensors(const aqo_defs* D, const aqo_ll* L, const int* imove,
                 const float* r, const float* m, float* u, float* rho,
                 float* p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 0)
            continue;
        const float* r_i = r + (size_t)i * vs;
        float ua[3] = { 0.f, 0.f, 0.f }, ra = 0.f, pa = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (i == j)
                continue;
            if (imove[j] != 1)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float rho_j = rho[j], m_j = m[j], p_j = p[j];
            const float w_ij = kernelW(q, dims) * D->CONW * m_j / rho_j;
            for (int d = 0; d < dims; d++)
                ua[d] += u[(size_t)j * vs + d] * w_ij;
            ra += rho_j * w_ij;
            pa += p_j * w_ij;
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            u[(size_t)i * vs + d] = ua[d];
        rho[i] = ra;
        p[i] = pa;
    }
}

/* cfd/SensorsRenormalization.cl:42-67 (whole vec divided) */
void aqo_sensors_renorm(const int* imove, const float* shepard, float* u,
                        float* rho, float* p, aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 0)
            continue;
        float s = shepard[i];
        if (s < 1.0E-6f)
            s = 1.f;
        for (int c = 0; c < vs; c++)
            u[(size_t)i * vs + c] /= s;
        rho[i] /= s;
        p[i] /= s;
    }
}

/* ------------------------------ delta-SPH -------------------------------- */
/* basic/deltaSPH.cl:57-71 */
void aqo_dsph_simple(const aqo_usize* iset, const int* imove, float* lap_p_corr,
                     const float* refd, aqo_usize N, const float* g, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        for (int c = 0; c < vs; c++)
            lap_p_corr[(size_t)i * vs + c] = refd[iset[i]] * g[c];
    }
}

/* basic/deltaSPH.cl:94-145 */
void aqo_dsph_full(const aqo_defs* D, const aqo_ll* L, const int* imove,
                   const float* r, const float* rho, const float* m,
                   const float* p, float* lap_p_corr)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float p_i = p[i];
        float gp[3] = { 0.f, 0.f, 0.f };
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((i == j) || (imove[j] != 1))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float f_ij = kernelF(q, dims) * D->CONF * m[j] / rho[j];
            const float c = (p[j] - p_i) * f_ij;
            for (int d = 0; d < dims; d++)
                gp[d] += c * r_ij[d];
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            lap_p_corr[(size_t)i * vs + d] = gp[d];
    }
}

/* basic/deltaSPH.cl:160-173; MATRIX_DOT types/3D.h:225-229, 2D.h:199-201 */
void aqo_dsph_full_mls(const int* imove, const float* mls, float* lap_p_corr,
                       aqo_usize N, int dims)
{
    const int vs = VS(dims), ms = MS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        const float* M = mls + (size_t)i * ms;
        float* v = lap_p_corr + (size_t)i * vs;
        if (dims == 3) {
            const float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
            const float y = M[4] * v[0] + M[5] * v[1] + M[6] * v[2];
            const float z = M[8] * v[0] + M[9] * v[1] + M[10] * v[2];
            v[0] = x;
            v[1] = y;
            v[2] = z;
            v[3] = 0.f;
        } else {
            const float x = M[0] * v[0] + M[1] * v[1];
            const float y = M[2] * v[0] + M[3] * v[1];
            v[0] = x;
            v[1] = y;
        }
    }
}

/* basic/deltaSPH.cl:191-242 */
void aqo_dsph_lapp(const aqo_defs* D, const aqo_ll* L, const int* imove,
                   const float* r, const float* rho, const float* m,
                   const float* p, float* lap_p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float p_i = p[i];
        float lp = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((i == j) || (imove[j] != 1))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float f_ij = kernelF(q, dims) * D->CONF * m[j] / rho[j];
            lp += (p[j] - p_i) * f_ij;
        }
        NEIGHS_END
        lap_p[i] = lp;
    }
}

/* basic/deltaSPH.cl:261-313: starts from the old lap_p[i] (:287) */
void aqo_dsph_lapp_corr(const aqo_defs* D, const aqo_ll* L, const int* imove,
                        const float* r, const float* rho, const float* m,
                        const float* lap_p_corr, float* lap_p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* g_i = lap_p_corr + (size_t)i * vs;
        float lp = lap_p[i];
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((i == j) || (imove[j] != 1))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            float g_ij[3];
            for (int d = 0; d < dims; d++)
                g_ij[d] = lap_p_corr[(size_t)j * vs + d] + g_i[d];
            const float f_ij = kernelF(q, dims) * D->CONF * m[j] / rho[j];
            lp -= 0.5f * dotv(g_ij, r_ij, dims) * f_ij;
        }
        NEIGHS_END
        lap_p[i] = lp;
    }
}

/* basic/deltaSPH.cl:330-350 */
void aqo_dsph_apply(const aqo_usize* iset, const int* imove, const float* rho,
                    const float* lap_p, float* drhodt, const float* refd,
                    const float* delta, aqo_usize N, float dt)
{
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        const aqo_usize s = iset[i];
        const float delta_f = delta[s] * dt * rho[i] / refd[s];
        drhodt[i] += delta_f * lap_p[i];
    }
}

/* --------------------------------- MLS ----------------------------------- */
/* basic/MLS.cl:58-112; outer(): types/3D.h:290-297, 2D.h:241-247.  imove is
 * compared with the unsigned mls_imove (int -> uint conversion). */
void aqo_mls(const aqo_defs* D, const aqo_ll* L, const int* imove,
             const float* r, const float* rho, const float* m, float* mls,
             aqo_usize mls_imove)
{
    const int dims = D->dims, vs = VS(dims), ms = MS(dims);
    const int rs = (dims == 3) ? 4 : 2; /* row stride of the matrix */
    AQO_FOR_I(L->N) {
        if ((aqo_usize)imove[i] != mls_imove)
            continue;
        const float* r_i = r + (size_t)i * vs;
        float M[16];
        for (int k = 0; k < 16; k++)
            M[k] = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (i == j)
                continue;
            if ((aqo_usize)imove[j] != mls_imove)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float f_ij = kernelF(q, dims) * D->CONF * m[j] / rho[j];
            for (int a = 0; a < dims; a++)
                for (int b = 0; b < dims; b++)
                    M[a * rs + b] += r_ij[a] * (f_ij * r_ij[b]);
        }
        NEIGHS_END
        memcpy(mls + (size_t)i * ms, M, sizeof(float) * ms);
    }
}

/* types/3D.h:300-341 (det, inv, MATRIX_INV), 2D.h:250-282 */
static void mat3_mul(const float* A, const float* B, float* C)
{
    /* MATRIX_MUL, 3D.h:241-245: 3x3 block, last row/col zero */
    for (int k = 0; k < 16; k++)
        C[k] = 0.f;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            C[a * 4 + b] = A[a * 4 + 0] * B[0 * 4 + b] +
                           A[a * 4 + 1] * B[1 * 4 + b] +
                           A[a * 4 + 2] * B[2 * 4 + b];
}
static void mat3_T(const float* A, float* T)
{
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++)
            T[a * 4 + b] = A[b * 4 + a];
}
static void mat3_inv(const float* m, float* o)
{
    const float det = m[0] * (m[5] * m[10] - m[6] * m[9]) +
                      m[1] * (m[6] * m[8] - m[4] * m[10]) +
                      m[2] * (m[4] * m[9] - m[5] * m[8]);
    const float d = 1.f / det;
    for (int k = 0; k < 16; k++)
        o[k] = 0.f;
    if (fabsf(d) > 1.e16f) {
        o[0] = o[5] = o[10] = o[15] = 1.f; /* MAT_ALL_EYE */
        return;
    }
    o[0] = (m[5] * m[10] - m[6] * m[9]) * d;
    o[1] = (m[2] * m[9] - m[1] * m[10]) * d;
    o[2] = (m[1] * m[6] - m[2] * m[5]) * d;
    o[4] = (m[6] * m[8] - m[4] * m[10]) * d;
    o[5] = (m[0] * m[10] - m[2] * m[8]) * d;
    o[6] = (m[2] * m[4] - m[0] * m[6]) * d;
    o[8] = (m[4] * m[9] - m[5] * m[8]) * d;
    o[9] = (m[1] * m[8] - m[0] * m[9]) * d;
    o[10] = (m[0] * m[5] - m[1] * m[4]) * d;
    o[15] = 1.f;
}

/* basic/MLS.cl:130-143 */
void aqo_mls_inv(const int* imove, float* mls, aqo_usize N,
                 aqo_usize mls_imove, int dims)
{
    AQO_FOR_I(N) {
        if ((aqo_usize)imove[i] != mls_imove)
            continue;
        if (dims == 3) {
            float* M = mls + 16 * (size_t)i;
            float T[16], TM[16], I[16], R[16];
            mat3_T(M, T);
            mat3_mul(T, M, TM);
            mat3_inv(TM, I);
            mat3_mul(I, T, R);
            memcpy(M, R, sizeof(R));
        } else {
            float* M = mls + 4 * (size_t)i;
            /* T = M.s0213 ; TM = MATRIX_MUL(T, M) (2D.h:205-208) */
            const float T[4] = { M[0], M[2], M[1], M[3] };
            float TM[4] = { T[0] * M[0] + T[1] * M[2], T[0] * M[1] + T[1] * M[3],
                            T[2] * M[0] + T[3] * M[2], T[2] * M[1] + T[3] * M[3] };
            const float d = 1.f / (TM[0] * TM[3] - TM[1] * TM[2]);
            float I[4];
            if (fabsf(d) > 1.e16f) {
                /* 2D.h:274 returns MAT_ALL_EYE, which 2D.h never defines: the
                 * reference would not compile that branch in 2-D; the
                 * Reduction header's MAT_EYE (1,0,0,1) is the evident intent */
                I[0] = 1.f; I[1] = 0.f; I[2] = 0.f; I[3] = 1.f;
            } else {
                I[0] = TM[3] * d; I[1] = -TM[1] * d;
                I[2] = -TM[2] * d; I[3] = TM[0] * d;
            }
            const float R[4] = { I[0] * T[0] + I[1] * T[2], I[0] * T[1] + I[1] * T[3],
                                 I[2] * T[0] + I[3] * T[2], I[2] * T[1] + I[3] * T[3] };
            memcpy(M, R, sizeof(R));
        }
    }
}

/* ------------------------------- BIe ------------------------------------- */
/* cfd/Boundary/BIe/Interactions.cl:48-108 */
void aqo_bie_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove,
                          const float* r, const float* normal, const float* u,
                          const float* m, float* grad_w_bi, float* div_u_bi)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        float gw[3] = { 0.f, 0.f, 0.f }, du = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (imove[j] != -3)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float* n_j = normal + (size_t)j * vs;
            const float* u_j = u + (size_t)j * vs;
            const float area_j = m[j];
            const float w = kernelW(q, dims) * D->CONW * area_j;
            float grad_w[3];
            for (int d = 0; d < dims; d++) {
                grad_w[d] = n_j[d] * w;
                gw[d] += grad_w[d];
            }
            du -= dotv(u_j, grad_w, dims);
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            grad_w_bi[(size_t)i * vs + d] = gw[d];
        div_u_bi[i] = du;
    }
}

/* cfd/Boundary/BIe/Interactions.cl:124-170 */
void aqo_bie_p_boundary(const aqo_defs* D, const aqo_ll* L, const int* imove,
                        const float* r, const float* m, const float* rho,
                        float* p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != -3)
            continue;
        const float* r_i = r + (size_t)i * vs;
        float pa = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (imove[j] != 1)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            pa += 2.f * p[j] * kernelW(q, dims) * D->CONW * m[j] / rho[j];
        }
        NEIGHS_END
        p[i] = pa;
    }
}

/* cfd/Boundary/BIe/Rates.cl:44-62 (whole vec arithmetic) */
void aqo_bie_rates(const int* imove, const float* rho, const float* p,
                   const float* u, const float* grad_w_bi,
                   const float* div_u_bi, float* grad_p, float* div_u,
                   aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        const float f = 2.f * p[i] / rho[i];
        float d = 0.f;
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            grad_p[k] += f * grad_w_bi[k];
            const float t = u[k] * grad_w_bi[k];
            d = (c == 0) ? t : d + t;
        }
        div_u[i] -= 2.f * rho[i] * (d + div_u_bi[i]);
    }
}

/* cfd/Boundary/BIe/Rates.cl:76-91 */
void aqo_bie_filter_press(const aqo_usize* iset, const int* imove, float* p,
                          aqo_usize forces_iset, aqo_usize N)
{
    AQO_FOR_I(N) {
        if (imove[i] != -3)
            continue;
        if (iset[i] != forces_iset)
            p[i] = 0.f;
    }
}

/* cfd/Boundary/BIe/Rates.cl:109-135; moment_p is always a vec4 */
void aqo_bie_force_press(const int* imove, const float* r, const float* normal,
                         const float* m, const float* p, float* force_p,
                         float* moment_p, const float* forces_r, aqo_usize N,
                         int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        float* fo = force_p + (size_t)i * vs;
        float* mo = moment_p + (size_t)i * 4;
        if (imove[i] != -3) {
            for (int c = 0; c < vs; c++)
                fo[c] = 0.f;
            mo[0] = mo[1] = mo[2] = mo[3] = 0.f;
            continue;
        }
        float F[4] = { 0.f, 0.f, 0.f, 0.f }, R[4] = { 0.f, 0.f, 0.f, 0.f };
        for (int d = 0; d < dims; d++) {
            F[d] = p[i] * m[i] * normal[(size_t)i * vs + d];
            R[d] = r[(size_t)i * vs + d] - forces_r[d];
            fo[d] = F[d];
        }
        mo[0] = R[1] * F[2] - R[2] * F[1];
        mo[1] = R[2] * F[0] - R[0] * F[2];
        mo[2] = R[0] * F[1] - R[1] * F[0];
        mo[3] = 0.f;
    }
}

/* cfd/Boundary/BIe/ElasticBounce.cl:64-152; dr_factor = __DR_FACTOR__ (default 0.5f, :31-33),
 * __MIN_BOUND_DIST__ = 0.0f (:34-36).  Order dependent: state is mutated
 * inside the neighbour loop. */
void aqo_bie_elastic_bounce(const aqo_ll* L, const int* imove,
                            const float* r_in, const float* normal,
                            const float* m, const float* u_in, float* dudt,
                            float dt, float dr_factor, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        if (!dt)
            continue;
        const float* r_i = r_in + (size_t)i * vs;
        const float* ui = u_in + (size_t)i * vs;
        float dudt_l[3], U[3];
        for (int d = 0; d < dims; d++) {
            dudt_l[d] = dudt[(size_t)i * vs + d];
            U[d] = ui[d] + 0.5f * dt * dudt_l[d];
        }
        NEIGHS_BEGIN(L, i, dims)
        {
            if (imove[j] != -3)
                continue;
            float r_ij[3];
            const float* n_j = normal + (size_t)j * vs;
            for (int d = 0; d < dims; d++)
                r_ij[d] = r_in[(size_t)j * vs + d] - r_i[d];
            const float rn = dotv(r_ij, n_j, dims);
            if (rn < 0.f)
                continue;
            const float dr = (dims == 3) ? sqrtf(m[j]) : m[j];
            const float R = dr_factor * dr;
            float rt[3];
            for (int d = 0; d < dims; d++)
                rt[d] = r_ij[d] - rn * n_j[d];
            if (dotv(rt, rt, dims) >= R * R)
                continue;
            const float drn = dt * dotv(U, n_j, dims);
            if (drn < 0.f)
                continue;
            if (rn - drn <= 0.0f * dr) {
                float uu[3], u_r[3];
                for (int d = 0; d < dims; d++)
                    uu[d] = ui[d] + dt * dudt_l[d];
                const float un = dotv(uu, n_j, dims);
                for (int d = 0; d < dims; d++)
                    u_r[d] = uu[d] - 2.f * un * n_j[d];
                for (int d = 0; d < dims; d++) {
                    dudt_l[d] = (u_r[d] - ui[d]) / dt;
                    U[d] = ui[d] + 0.5f * dt * dudt_l[d];
                }
            }
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            dudt[(size_t)i * vs + d] = dudt_l[d];
    }
}

/* cfd/Boundary/BIe/ElasticBounce.cl:168-184 */
void aqo_bie_force_bound(const int* imove, const float* m,
                         const float* dudt_preelastic,
                         const float* dudt_elastic, float* force_elastic,
                         aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N)
        for (int c = 0; c < vs; c++) {
            const size_t k = (size_t)i * vs + c;
            force_elastic[k] = (imove[i] != 1) ? 0.f
                : -m[i] * (dudt_elastic[k] - dudt_preelastic[k]);
        }
}

/* cfd/Boundary/BIe/PST.cl:62-110.  DIMS is the evaluated define (a float
 * literal, basic.xml:119).  Order dependent: r[i] moves inside the loop and is
 * re-read for the next element. */
void aqo_bie_pst(const aqo_ll* L, const int* imove, float* r,
                 const float* normal, const float* m, const float* rho,
                 float DIMS_define, float dr_factor, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float Ri = 0.5f * powf(m[i] / rho[i], 1.f / DIMS_define);
        float* r_i = r + (size_t)i * vs;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (imove[j] != -3)
                continue;
            float r_ij[3];
            const float* n_j = normal + (size_t)j * vs;
            for (int d = 0; d < dims; d++)
                r_ij[d] = r[(size_t)j * vs + d] - r_i[d];
            const float rn = dotv(r_ij, n_j, dims);
            if (fabsf(rn) > Ri)
                continue;
            const float dr = (dims == 3) ? sqrtf(m[j]) : m[j];
            const float Rj = dr_factor * dr;
            float rt[3];
            for (int d = 0; d < dims; d++)
                rt[d] = r_ij[d] - rn * n_j[d];
            if (dotv(rt, rt, dims) >= Rj * Rj)
                continue;
            for (int d = 0; d < dims; d++)
                r_i[d] += (rn - Ri) * n_j[d];
        }
        NEIGHS_END
    }
}

/* ============================ prescribed motions ========================== *
 * cfd/Motions/{Transform,UnTransform,Velocity,Acceleration}.cl (preset
 * resources/Presets/src/cfd/motion.xml:55-72): rigid motion of the non-fluid
 * particles of one set, Euler-XYZ angles (phi, theta, psi) = motion_a.xyz.
 * vec scalars (motion_r, ...) are arrays of VS(dims) floats, vec4 ones of 4. */

/* rotate (x, y, z) of up to three vectors in place: along x, then y (3-D only), then z;
 * (s_phi, s_theta, s_psi) are passed signed so that UnTransform can reuse the stages */
static inline void rot_x(float* v, float c, float s)
{
    const float y = v[1], z = v[2];
    v[1] = c * y - s * z;
    v[2] = s * y + c * z;
}
static inline void rot_y(float* v, float c, float s)
{
    const float x = v[0], z = v[2];
    v[0] = c * x + s * z;
    v[2] = -s * x + c * z;
}
static inline void rot_z(float* v, float c, float s)
{
    const float x = v[0], y = v[1];
    v[0] = c * x - s * y;
    v[1] = s * x + c * y;
}
/* OpenCL normalize over the whole vec (w = 0 in 3-D): v / length(v), products summed left to right */
static inline void normalize_vec(float* v, int vs)
{
    float d = v[0] * v[0];
    for (int k = 1; k < vs; k++)
        d = d + v[k] * v[k];
    const float l = sqrtf(d);
    for (int k = 0; k < vs; k++)
        v[k] = v[k] / l;
}

/* cfd/Motions/Transform.cl:68-134 */
void aqo_motion_transform(const unsigned* iset, const int* imove, float* r, float* normal,
                          float* tangent, aqo_usize N, unsigned motion_iset,
                          const float* motion_r, const float* motion_a, int dims)
{
    const int vs = VS(dims);
    const float cphi = cosf(motion_a[0]), sphi = sinf(motion_a[0]);
    const float cth = cosf(motion_a[1]), sth = sinf(motion_a[1]);
    const float cpsi = cosf(motion_a[2]), spsi = sinf(motion_a[2]);
    AQO_FOR_I(N) {
        if (iset[i] != motion_iset || imove[i] == 1)
            continue;
        float* v[3] = { r + (size_t)i * vs, normal + (size_t)i * vs, tangent + (size_t)i * vs };
        for (int a = 0; a < 3; a++) {
            if (dims == 3) {
                rot_x(v[a], cphi, sphi);
                rot_y(v[a], cth, sth);
            }
            rot_z(v[a], cpsi, spsi);
        }
        for (int k = 0; k < vs; k++)
            v[0][k] = v[0][k] + motion_r[k];
        normalize_vec(v[1], vs);
        normalize_vec(v[2], vs);
    }
}

/* cfd/Motions/UnTransform.cl:54-119: the inverse rotations in the inverse order */
void aqo_motion_untransform(const unsigned* iset, const int* imove, float* r, float* normal,
                            float* tangent, aqo_usize N, unsigned motion_iset,
                            const float* motion_r_in, const float* motion_a_in, int dims)
{
    const int vs = VS(dims);
    const float cphi = cosf(motion_a_in[0]), sphi = -sinf(motion_a_in[0]);
    const float cth = cosf(motion_a_in[1]), sth = -sinf(motion_a_in[1]);
    const float cpsi = cosf(motion_a_in[2]), spsi = -sinf(motion_a_in[2]);
    AQO_FOR_I(N) {
        if (iset[i] != motion_iset || imove[i] == 1)
            continue;
        float* v[3] = { r + (size_t)i * vs, normal + (size_t)i * vs, tangent + (size_t)i * vs };
        for (int k = 0; k < vs; k++)
            v[0][k] = v[0][k] - motion_r_in[k];
        for (int a = 0; a < 3; a++) {
            rot_z(v[a], cpsi, spsi);
            if (dims == 3) {
                rot_y(v[a], cth, sth);
                rot_x(v[a], cphi, sphi);
            }
        }
    }
}

/* cfd/Motions/Velocity.cl:74-121 and Acceleration.cl (same shape): omega x r in the local
 * frame (r is still untransformed there), rotated, plus the linear part */
static void motion_rate(const unsigned* iset, const int* imove, const float* r, float* out,
                        aqo_usize N, unsigned motion_iset, const float* lin,
                        const float* motion_a, const float* w, int dims)
{
    const int vs = VS(dims);
    const float cphi = cosf(motion_a[0]), sphi = sinf(motion_a[0]);
    const float cth = cosf(motion_a[1]), sth = sinf(motion_a[1]);
    const float cpsi = cosf(motion_a[2]), spsi = sinf(motion_a[2]);
    AQO_FOR_I(N) {
        if (iset[i] != motion_iset || imove[i] == 1)
            continue;
        const float* p = r + (size_t)i * vs;
        float v[4] = { 0.f, 0.f, 0.f, 0.f };
        if (dims == 2) {
            v[0] = -w[2] * p[1];
            v[1] = w[2] * p[0];
        } else { /* cross(float4, float4): w = 0 */
            v[0] = w[1] * p[2] - w[2] * p[1];
            v[1] = w[2] * p[0] - w[0] * p[2];
            v[2] = w[0] * p[1] - w[1] * p[0];
            rot_x(v, cphi, sphi);
            rot_y(v, cth, sth);
        }
        rot_z(v, cpsi, spsi);
        for (int k = 0; k < vs; k++)
            out[(size_t)i * vs + k] = v[k] + lin[k];
    }
}

void aqo_motion_velocity(const unsigned* iset, const int* imove, const float* r, float* u,
                         aqo_usize N, unsigned motion_iset, const float* motion_drdt,
                         const float* motion_a, const float* motion_dadt, int dims)
{
    motion_rate(iset, imove, r, u, N, motion_iset, motion_drdt, motion_a, motion_dadt, dims);
}

/* Acceleration.cl declares motion_r too and does not use it */
void aqo_motion_acceleration(const unsigned* iset, const int* imove, const float* r, float* dudt,
                             aqo_usize N, unsigned motion_iset, const float* motion_ddrddt,
                             const float* motion_a, const float* motion_ddaddt, int dims)
{
    motion_rate(iset, imove, r, dudt, N, motion_iset, motion_ddrddt, motion_a, motion_ddaddt, dims);
}

/* ============================== energy report ============================= *
 * cfd/Energy/Energy.cl (preset resources/Presets/src/cfd/energy.xml): per-particle
 * power and energy terms, summed afterwards by reduction tools. */
static inline float dot_vec(const float* a, const float* b, int vs)
{
    float d = a[0] * b[0];
    for (int k = 1; k < vs; k++)
        d = d + a[k] * b[k];
    return d;
}

/* Energy.cl:59-87 */
void aqo_energy_power(float* energy_dekdt, float* energy_depdt, float* energy_decdt, const int* imove,
                      const float* u, const float* rho, const float* m, const float* p,
                      const float* dudt, const float* drhodt, aqo_usize N, const float* g, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1) {
            energy_dekdt[i] = 0.f;
            energy_depdt[i] = 0.f;
            energy_decdt[i] = 0.f;
            continue;
        }
        const float* ui = u + (size_t)i * vs;
        energy_depdt[i] = -m[i] * dot_vec(g, ui, vs);
        energy_dekdt[i] = m[i] * dot_vec(ui, dudt + (size_t)i * vs, vs);
        energy_decdt[i] = m[i] * p[i] / (rho[i] * rho[i]) * drhodt[i];
    }
}

/* Energy.cl:114-145 */
void aqo_energy_energy(float* energy_ek, float* energy_ep, float* energy_ec, const unsigned* iset,
                       const int* imove, const float* r, const float* u, const float* rho,
                       const float* m, const float* refd, aqo_usize N, const float* g, float cs,
                       int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1) {
            energy_ek[i] = 0.f;
            energy_ep[i] = 0.f;
            energy_ec[i] = 0.f;
            continue;
        }
        const float* ui = u + (size_t)i * vs;
        energy_ek[i] = 0.5f * m[i] * dot_vec(ui, ui, vs);
        energy_ep[i] = -m[i] * dot_vec(g, r + (size_t)i * vs, vs);
        const float rho0 = refd[iset[i]];
        energy_ec[i] = m[i] * cs * cs * (rho0 / rho[i] + logf(rho[i] / rho0) - 1.f);
    }
}

/* ===================== small presets next to the hot path ================== */
/* cfd/Energy/EnergyKin.cl:38-54 */
void aqo_energy_kin(float* energy_kin, const int* imove, const float* u, const float* m, aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1) {
            energy_kin[i] = 0.f;
            continue;
        }
        const float* ui = u + (size_t)i * vs;
        energy_kin[i] = 0.5f * m[i] * dot_vec(ui, ui, vs);
    }
}

/* cfd/Forces/Forces.cl:50-84 */
void aqo_forces(float* forces_f, float* forces_m, const int* imove, const float* r, const float* dudt,
                const float* m, aqo_usize N, const float* g, const float* forces_r, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        float* f = forces_f + (size_t)i * vs;
        float* mo = forces_m + (size_t)i * 4;
        if (imove[i] != 1) {
            for (int k = 0; k < vs; k++)
                f[k] = 0.f;
            mo[0] = mo[1] = mo[2] = mo[3] = 0.f;
            continue;
        }
        float arm[4] = { 0.f, 0.f, 0.f, 0.f }, acc[4] = { 0.f, 0.f, 0.f, 0.f };
        for (int k = 0; k < vs; k++) {
            arm[k] = r[(size_t)i * vs + k] - forces_r[k];
            acc[k] = g[k] - dudt[(size_t)i * vs + k];
        }
        const float mass = m[i];
        for (int k = 0; k < vs; k++)
            f[k] = mass * acc[k];
        mo[2] = mass * (arm[0] * acc[1] - arm[1] * acc[0]);
        mo[3] = 0.f;
        if (dims == 3) {
            mo[0] = mass * (arm[1] * acc[2] - arm[2] * acc[1]);
            mo[1] = mass * (arm[2] * acc[0] - arm[0] * acc[2]);
        } else {
            mo[0] = 0.f;
            mo[1] = 0.f;
        }
    }
}

/* basic/DensityClamp.cl:41-52 */
void aqo_density_clamp(float* rho_in, aqo_usize N, float rho_min, float rho_max)
{
    AQO_FOR_I(N) {
        if (rho_in[i] < rho_min)
            rho_in[i] = rho_min;
        if (rho_in[i] > rho_max)
            rho_in[i] = rho_max;
    }
}

/* basic/IdInverse.cl:33-42 */
void aqo_id_inverse(const aqo_usize* id, aqo_usize* id_inverse, aqo_usize N)
{
    AQO_FOR_I(N)
        id_inverse[id[i]] = i;
}

/* ================= basic/time_scheme/adam_bashforth.cl ===================== *
 * The four history levels travel as arrays of 4 pointers (as1..as4). */
/* :122-146 */
void aqo_ab_sort(const float* const* dudt_as_in, float* const* dudt_as, const float* const* drhodt_as_in,
                 float* const* drhodt_as, const aqo_usize* id_sorted, aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        const size_t o = id_sorted[i];
        for (int l = 0; l < 4; l++) {
            for (int k = 0; k < vs; k++)
                dudt_as[l][o * vs + k] = dudt_as_in[l][(size_t)i * vs + k];
            drhodt_as[l][o] = drhodt_as_in[l][i];
        }
    }
}

/* the extrapolated rate, DYDT_1..DYDT_5 (:148-159): products and sums left to right in fp32 */
static inline float ab_rate(unsigned local_iter, float d0, float d1, float d2, float d3, float d4)
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

/* :207-255 */
void aqo_ab_corrector(const int* imove, float* r, float* u, const float* dudt, float* rho, const float* drhodt,
                      const float* const* dudt_as, const float* const* drhodt_as, aqo_usize N, float dt,
                      unsigned iter, unsigned steps, int dims)
{
    const int vs = VS(dims);
    const unsigned local_iter = iter < steps ? iter : steps;
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        for (int k = 0; k < vs; k++) {
            const size_t j = (size_t)i * vs + k;
            const float a = ab_rate(local_iter, dudt[j], dudt_as[0][j], dudt_as[1][j], dudt_as[2][j],
                                    dudt_as[3][j]);
            r[j] += dt * u[j] + 0.5f * dt * dt * a;
            u[j] += dt * a;
        }
        rho[i] += dt * ab_rate(local_iter, drhodt[i], drhodt_as[0][i], drhodt_as[1][i], drhodt_as[2][i],
                               drhodt_as[3][i]);
    }
}

/* :281-309 */
void aqo_ab_postcorrector(const float* const* dudt_as, const float* const* drhodt_as, const float* dudt,
                          const float* drhodt, float* const* dudt_as_in, float* const* drhodt_as_in,
                          aqo_usize N, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        for (int k = 0; k < vs; k++) {
            const size_t j = (size_t)i * vs + k;
            dudt_as_in[0][j] = dudt[j];
            dudt_as_in[1][j] = dudt_as[0][j];
            dudt_as_in[2][j] = dudt_as[1][j];
            dudt_as_in[3][j] = dudt_as[2][j];
        }
        drhodt_as_in[0][i] = drhodt[i];
        drhodt_as_in[1][i] = drhodt_as[0][i];
        drhodt_as_in[2][i] = drhodt_as[1][i];
        drhodt_as_in[3][i] = drhodt_as[2][i];
    }
}

/* cfd/Boundary/BI/NoSlip.cl:52-130: fluid i against the elements of the set noslip_iset, added to lap_u */
void aqo_bi_noslip(const aqo_defs* D, const aqo_ll* L, const unsigned* iset, const int* imove, const float* r,
                   const float* normal, const float* u, const float* rho, const float* m, float* lap_u,
                   unsigned noslip_iset, float dr)
{
    const int dims = D->dims, vs = VS(dims);
    const float cleary = dims == 3 ? 10.f : 8.f; /* :41-47 */
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* u_i = u + (size_t)i * vs;
        const float rho_i = rho[i];
        float* lap = lap_u + (size_t)i * vs;
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((imove[j] != -3) || (iset[j] != noslip_iset))
                continue;
            float r_ij[3] = { 0.f, 0.f, 0.f }, q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float* n_j = normal + (size_t)j * vs;
            const float area_j = m[j];
            const float w_ij = kernelW(q, dims) * D->CONW * area_j;
            float du[3] = { 0.f, 0.f, 0.f };
            for (int d = 0; d < dims; d++)
                du[d] = u[(size_t)j * vs + d] - u_i[d];
            const float r2 = (q * q + 0.01f) * D->H * D->H;
            const float c1 = cleary * w_ij * dotv(du, r_ij, dims) / (r2 * rho_i);
            for (int d = 0; d < dims; d++)
                lap[d] += c1 * n_j[d];
            const float dr_n = fmaxf(fabsf(dotv(r_ij, n_j, dims)), dr);
            const float dun = dotv(du, n_j, dims);
            const float c2 = 2.f * w_ij / (rho_i * dr_n);
            for (int d = 0; d < dims; d++)
                lap[d] += c2 * (du[d] - dun * n_j[d]);
        }
        NEIGHS_END
    }
}

/* ---------------------------------------------------------------------------
 * Remote (halo) terms of MLS and delta-SPH.  NOT reference kernels: the reference's MPI preset
 * (resources/Presets/src/cfd/MPI.xml) exchanges what cfd/Interactions.cl and the Shepard factor need
 * and nothing else, so its multi-process runs cannot use delta-SPH or MLS.  These are the j loops of
 * basic/MLS.cl:58-112 and basic/deltaSPH.cl:94-145, 191-242, 261-313 over the halo list, written the
 * way cfd/MPI.cl:328-485 writes its own (every halo particle counts, the result is ADDED to what the
 * local kernel left); cases_xml of the multi-device pipelines insert them next to their local twins.
 * L: icell/ihoc of the HALO list (mpi_icell, mpi_ihoc) with N = number of local particles;
 * icell_i: cells of the local particles. */
#define REMOTE_NEIGHS_BEGIN(L, icell_i, i, dims)                               \
    {                                                                          \
        const aqo_usize c_i__ = (icell_i)[i];                                  \
        const aqo_usize nx__ = (L)->ncells[0], ny__ = (L)->ncells[1];          \
        const int kz__ = ((dims) == 3) ? 1 : 0;                                \
        for (int ci__ = -1; ci__ <= 1; ci__++)                                 \
            for (int cj__ = -1; cj__ <= 1; cj__++)                             \
                for (int ck__ = -kz__; ck__ <= kz__; ck__++) {                 \
                    const aqo_usize c_j__ =                                    \
                        c_i__ + (aqo_usize)ci__ + (aqo_usize)cj__ * nx__ +     \
                        (aqo_usize)ck__ * nx__ * ny__;                         \
                    for (aqo_usize j = (L)->ihoc[c_j__];                       \
                         (j < (L)->N) && ((L)->icell[j] == c_j__); j++) {

void aqo_mpi_mls(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i, const int* imove,
                 const float* r, const float* mpi_r, const float* mpi_rho, const float* mpi_m,
                 float* mls, aqo_usize mls_imove)
{
    const int dims = D->dims, vs = VS(dims), ms = MS(dims);
    const int rs = (dims == 3) ? 4 : 2;
    AQO_FOR_I(L->N) {
        if ((aqo_usize)imove[i] != mls_imove)
            continue;
        const float* r_i = r + (size_t)i * vs;
        float M[16];
        for (int k = 0; k < 16; k++)
            M[k] = 0.f;
        REMOTE_NEIGHS_BEGIN(L, icell_i, i, dims)
        {
            float r_ij[3], q;
            if (!pair_q(D, r_i, mpi_r + (size_t)j * vs, r_ij, &q))
                continue;
            const float f_ij = kernelF(q, dims) * D->CONF * mpi_m[j] / mpi_rho[j];
            for (int a = 0; a < dims; a++)
                for (int b = 0; b < dims; b++)
                    M[a * rs + b] += r_ij[a] * (f_ij * r_ij[b]);
        }
        NEIGHS_END
        for (int k = 0; k < ms; k++)
            mls[(size_t)i * ms + k] += M[k];
    }
}

void aqo_mpi_dsph_full_lapp(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i,
                            const int* imove, const float* r, const float* p, const float* mpi_r,
                            const float* mpi_rho, const float* mpi_m, const float* mpi_p,
                            float* lap_p_corr, float* lap_p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float p_i = p[i];
        float gp[3] = { 0.f, 0.f, 0.f }, lp = 0.f;
        REMOTE_NEIGHS_BEGIN(L, icell_i, i, dims)
        {
            float r_ij[3], q;
            if (!pair_q(D, r_i, mpi_r + (size_t)j * vs, r_ij, &q))
                continue;
            const float f_ij = kernelF(q, dims) * D->CONF * mpi_m[j] / mpi_rho[j];
            const float c = (mpi_p[j] - p_i) * f_ij;
            for (int d = 0; d < dims; d++)
                gp[d] += c * r_ij[d];
            lp += c;
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            lap_p_corr[(size_t)i * vs + d] += gp[d];
        lap_p[i] += lp;
    }
}

void aqo_mpi_dsph_lapp_corr(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i,
                            const int* imove, const float* r, const float* lap_p_corr,
                            const float* mpi_r, const float* mpi_rho, const float* mpi_m,
                            const float* mpi_lap_p_corr, float* lap_p)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* g_i = lap_p_corr + (size_t)i * vs;
        float acc = 0.f;
        REMOTE_NEIGHS_BEGIN(L, icell_i, i, dims)
        {
            float r_ij[3], q;
            if (!pair_q(D, r_i, mpi_r + (size_t)j * vs, r_ij, &q))
                continue;
            float g_ij[3];
            for (int d = 0; d < dims; d++)
                g_ij[d] = mpi_lap_p_corr[(size_t)j * vs + d] + g_i[d];
            const float f_ij = kernelF(q, dims) * D->CONF * mpi_m[j] / mpi_rho[j];
            acc += dotv(g_ij, r_ij, dims) * f_ij;
        }
        NEIGHS_END
        lap_p[i] -= 0.5f * acc;
    }
}

/* ---------------------------------------------------------------------------
 * cfd/Boundary/Symmetry/Mirror.cl (preset cfd/symmetry.xml): an infinite symmetry plane made of
 * mirrored copies of the particles within the kernel support of it, taken from the buffer rows. */
static float aqo_dotv(const float* a, const float* b, int n)
{
    float s = a[0] * b[0];
    for (int k = 1; k < n; k++)
        s += a[k] * b[k];
    return s;
}

/* Mirror.cl:37-55 */
void aqo_sym_drop(int* imove, float* r, aqo_usize N, const float* symmetry_r, const float* symmetry_n,
                  const float* domain_max, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] <= -255)
            continue;
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = r[(size_t)i * vs + k] - symmetry_r[k];
        if (aqo_dotv(d, symmetry_n, vs) >= 0.f) {
            for (int k = 0; k < vs; k++) /* VEC_ONE: w = 0 in 3-D (types/3D.h:38) */
                r[(size_t)i * vs + k] = domain_max[k] + ((k < dims) ? 1.f : 0.f);
            imove[i] = -256;
        }
    }
}

/* Mirror.cl:72-96 */
void aqo_sym_detect(const aqo_defs* D, const int* imove, const float* r_in, aqo_usize* imirror, aqo_usize N,
                    const float* symmetry_r, const float* symmetry_n)
{
    const int vs = VS(D->dims);
    AQO_FOR_I(N) {
        if (imove[i] <= -255) {
            imirror[i] = 0;
            continue;
        }
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = symmetry_r[k] - r_in[(size_t)i * vs + k];
        imirror[i] = (fabsf(aqo_dotv(d, symmetry_n, vs)) <= D->SUPPORT * D->H) ? 1u : 0u;
    }
}

/* Mirror.cl:114-117: v = -2 (u . n) n over the XYZ components; out = base + v */
static void aqo_reflect_add(const float* base, const float* u, const float* n, int dims, float* out)
{
    const float f = -2.f * aqo_dotv(u, n, dims);
    for (int k = 0; k < dims; k++)
        out[k] = base[k] + f * n[k];
}

/* Mirror.cl:140-186 (a sequential loop: the script's only cross-row access, imove of a buffer row
 * that another work-item is overwriting, cannot change what that row's work-item does -- its sorted
 * imirror entry is 0 either way) */
void aqo_sym_feed(int* imove, int* iset, const aqo_usize* imirror, const aqo_usize* imirror_invperm,
                  aqo_usize* mirror_src, float* normal, float* tangent, float* r_in, aqo_usize N,
                  aqo_usize nbuffer, const float* symmetry_r, const float* symmetry_n, int dims)
{
    const int vs = VS(dims);
    for (aqo_usize i = 0; i < N; i++) {
        if (imove[i] <= -255)
            continue;
        const aqo_usize j = imirror_invperm[i];
        if (imirror[j] != 1)
            continue;
        const aqo_usize i0 = N - nbuffer;
        const aqo_usize ii = i0 + (N - j - 1);
        if (ii >= N)
            continue;
        mirror_src[ii] = i;
        imove[ii] = imove[i];
        iset[ii] = iset[i];
        aqo_reflect_add(normal + (size_t)i * vs, normal + (size_t)i * vs, symmetry_n, dims, normal + (size_t)ii * vs);
        aqo_reflect_add(tangent + (size_t)i * vs, tangent + (size_t)i * vs, symmetry_n, dims,
                        tangent + (size_t)ii * vs);
        float rel[3];
        for (int k = 0; k < dims; k++)
            rel[k] = r_in[(size_t)i * vs + k] - symmetry_r[k];
        aqo_reflect_add(r_in + (size_t)i * vs, rel, symmetry_n, dims, r_in + (size_t)ii * vs);
    }
}

/* Mirror.cl:203-228 */
void aqo_sym_set(const aqo_usize* mirror_src, float* m, float* u_in, float* dudt_in, float* dudt, float* rho_in,
                 float* drhodt_in, float* drhodt, aqo_usize N, const float* symmetry_n, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        const aqo_usize ii = i, src = mirror_src[ii];
        if (src >= N)
            continue;
        m[ii] = m[src];
        rho_in[ii] = rho_in[src];
        drhodt[ii] = drhodt_in[ii] = drhodt_in[src];
        aqo_reflect_add(u_in + (size_t)src * vs, u_in + (size_t)src * vs, symmetry_n, dims, u_in + (size_t)ii * vs);
        float a[3];
        aqo_reflect_add(dudt_in + (size_t)src * vs, dudt_in + (size_t)src * vs, symmetry_n, dims, a);
        for (int k = 0; k < dims; k++)
            dudt[(size_t)ii * vs + k] = dudt_in[(size_t)ii * vs + k] = a[k];
    }
}

/* Mirror.cl:239-251 */
void aqo_sym_sort(const aqo_usize* mirror_src_in, aqo_usize* mirror_src, const aqo_usize* id_sorted, aqo_usize N)
{
    AQO_FOR_I(N)
        mirror_src[id_sorted[i]] = mirror_src_in[i];
}

/* ================= cfd/ideal_gas: the element-wise kernels ================= *
 * (the internal energy next to rho / u of the weakly compressible scheme: the presets under cfd/ideal_gas,
 * examples/2D/shock_*) */

/* cfd/ideal_gas/EOS.cl:56-70; EXCLUDED_PARTICLE :32-34 */
void aqo_ig_eos(const aqo_usize* iset, const int* imove, const float* rho, const float* eint, float* p,
                const float* gamma, aqo_usize N)
{
    AQO_FOR_I(N) {
        if ((imove[i] <= 0) && (imove[i] != -1))
            continue;
        p[i] = (gamma[iset[i]] - 1.0f) * rho[i] * eint[i];
    }
}

/* cfd/ideal_gas/Rates.cl:53-68 */
void aqo_ig_rates(const int* imove, const float* rho, const float* p, const float* div_u, float* deintdt,
                  aqo_usize N)
{
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        deintdt[i] = -p[i] / (rho[i] * rho[i]) * div_u[i];
    }
}

/* cfd/ideal_gas/Sort.cl:43-58 */
void aqo_ig_sort(const float* eint_in, float* eint, const float* deintdt, float* deintdt_in,
                 const aqo_usize* id_sorted, aqo_usize N)
{
    AQO_FOR_I(N) {
        const aqo_usize o = id_sorted[i];
        eint[o] = eint_in[i];
        deintdt_in[o] = deintdt[i];
    }
}

/* cfd/ideal_gas/TimeStep.cl:62-97; sound_speed.hcl:22-25; length() of a vec takes every component */
void aqo_ig_timestep(const aqo_defs* D, float* dt_var, const int* imove, const aqo_usize* iset, const float* u,
                     const float* rho, const float* p, aqo_usize N, float dt, float dt_min, float courant,
                     const float* div_u, const float* grad_p, const float* gamma)
{
    const int vs = VS(D->dims);
    AQO_FOR_I(N) {
        if (imove[i] <= 0) {
            dt_var[i] = dt;
            continue;
        }
        const float dxx = D->H;
        const float s_i = sqrtf(gamma[iset[i]] * p[i] / rho[i]);
        float g2 = 0.f, u2 = 0.f;
        for (int c = 0; c < vs; c++) {
            g2 = c ? g2 + grad_p[vs * i + c] * grad_p[vs * i + c] : grad_p[vs * i] * grad_p[vs * i];
            u2 = c ? u2 + u[vs * i + c] * u[vs * i + c] : u[vs * i] * u[vs * i];
        }
        const float lg = sqrtf(g2), lu = sqrtf(u2);
        const float a = 4.0f * dxx * div_u[i] / rho[i];
        const float dt_u1 = courant * 0.4f * dxx / sqrtf(a * a + s_i * s_i);
        const float dt_u2 = courant * sqrtf(dxx / lg);
        const float dt_u3 = courant * 0.4f * dxx / sqrtf(lu * lu + s_i * s_i);
        /* OpenCL min(x, y) = y < x ? y : x, max(x, y) = x < y ? y : x */
        float m12 = dt_u2 < dt_u1 ? dt_u2 : dt_u1;
        const float dt_u = dt_u3 < m12 ? dt_u3 : m12;
        const float lo = dt_u < dt ? dt_u : dt;
        dt_var[i] = lo < dt_min ? dt_min : lo;
    }
}

/* cfd/ideal_gas/riemann/Rates.cl:39-53 */
void aqo_ig_riemann_rates(const int* imove, const float* work_density, float* deintdt, aqo_usize N)
{
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        deintdt[i] = -work_density[i];
    }
}

/* cfd/ideal_gas/time_scheme/midpoint.cl: predictor :47-59, midpoint :75-88, relax :101-115, corrector :131-144 */
void aqo_ig_mp_predictor(const float* eint, const float* deintdt, float* eint_in, float* deintdt_in, aqo_usize N)
{
    AQO_FOR_I(N) {
        deintdt_in[i] = deintdt[i];
        eint_in[i] = eint[i];
    }
}
void aqo_ig_mp_midpoint(const int* imove, const float* eint_in, const float* deintdt, float* eint, aqo_usize N,
                        float dt)
{
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        eint[i] = eint_in[i] + 0.5f * dt * deintdt[i];
    }
}
void aqo_ig_mp_relax(const int* imove, const float* deintdt_in, float* deintdt, aqo_usize N, float relax_midpoint)
{
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        deintdt[i] = relax_midpoint * deintdt_in[i] + (1.f - relax_midpoint) * deintdt[i];
    }
}
void aqo_ig_mp_corrector(const int* imove, const float* eint_in, const float* deintdt, float* eint, aqo_usize N,
                         float dt)
{
    AQO_FOR_I(N) {
        if (imove[i] <= 0)
            continue;
        eint[i] = eint_in[i] + dt * deintdt[i];
    }
}

/* cfd/ideal_gas/symmetry/Mirror.cl:32-48 (a mirrored particle copies the energy of its source) */
void aqo_ig_sym_set(const aqo_usize* mirror_src, float* eint_in, float* deintdt_in, float* deintdt, aqo_usize N)
{
    AQO_FOR_I(N) {
        const aqo_usize s = mirror_src[i];
        if (s >= N)
            continue;
        eint_in[i] = eint_in[s];
        deintdt[i] = deintdt_in[i] = deintdt_in[s];
    }
}

/* cfd/ideal_gas/riemann/Interactions.cl:50-168: the pair terms of the acoustic Riemann solver between fluid
 * particles (LOCAL_MEM_SIZE is always defined, Kernel.cpp:411: the sums start at zero and overwrite) */
void aqo_ig_riemann_interactions(const aqo_defs* D, const aqo_ll* L, const aqo_usize* iset, const int* imove,
                                 const float* r, const float* u, const float* rho, const float* m, const float* p,
                                 float* grad_p, float* div_u, float* work_density, const float* gamma)
{
    const int dims = D->dims, vs = VS(dims);
    const float H = D->H;
    AQO_FOR_I(L->N) {
        if (imove[i] != 1)
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* u_i = u + (size_t)i * vs;
        const float p_i = p[i], rho_i = rho[i];
        const float s_i = sqrtf(gamma[iset[i]] * p_i / rho_i);
        const float rs_i = rho_i * s_i;
        float gp[3] = { 0.f, 0.f, 0.f }, du = 0.f, wd = 0.f;
        NEIGHS_BEGIN(L, i, dims)
        {
            if (i == j)
                continue;
            if (imove[j] != 1)
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float rho_j = rho[j], p_j = p[j], m_j = m[j];
            const float s_j = sqrtf(gamma[iset[j]] * p_j / rho_j);
            const float len = sqrtf(dotv(r_ij, r_ij, dims));
            float l_ij[3];
            for (int d = 0; d < dims; d++)
                l_ij[d] = r_ij[d] / len;
            const float u_R_i = dotv(u_i, l_ij, dims);
            const float u_R_j = dotv(u + (size_t)j * vs, l_ij, dims);
            const float rs_j = rho_j * s_j;
            const float auxiliary_val = 1.0f / (rs_j + rs_i);
            const float u_star = (u_R_j * rs_j + u_R_i * rs_i - p_j + p_i) * auxiliary_val;
            const float p_star = (p_j * rs_i + p_i * rs_j - rs_j * rs_i * (u_R_j - u_R_i)) * auxiliary_val;
            const float Wij_prima = -q * kernelF(q, dims) * D->CONW;
            const float aux = 2.0f * m_j / (rho_j * H) * (u_R_i - u_star) * Wij_prima;
            du += rho_i * aux;
            const float g = 2.0f * m_j * p_star / (rho_j * rho_i * H) * Wij_prima;
            for (int d = 0; d < dims; d++)
                gp[d] -= g * l_ij[d];
            wd += p_star / rho_i * aux;
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++)
            grad_p[(size_t)i * vs + d] = gp[d];
        work_density[i] = wd;
        div_u[i] = du;
    }
}

/* cfd/ideal_gas/time_scheme/euler.cl: predictor :44-56 (the same copies as midpoint.cl's), corrector :70-83;
 * improved_euler.cl: predictor :49-66, corrector :82-97 */
void aqo_ig_euler_corrector(const int* imove, float* eint, const float* deintdt, aqo_usize N, float dt)
{
    AQO_FOR_I(N) {
        if (imove[i] > 0)
            eint[i] += dt * deintdt[i];
    }
}
void aqo_ig_ie_predictor(const int* imove, const float* eint, const float* deintdt, float* eint_in,
                         float* deintdt_in, aqo_usize N, float dt)
{
    AQO_FOR_I(N) {
        float DT = dt;
        if (imove[i] <= 0)
            DT = 0.f;
        deintdt_in[i] = deintdt[i];
        eint_in[i] = eint[i] + DT * deintdt[i];
    }
}
void aqo_ig_ie_corrector(const int* imove, const float* deintdt, const float* deintdt_in, float* eint,
                         aqo_usize N, float dt)
{
    AQO_FOR_I(N) {
        if (imove[i] > 0) {
            const float DT = 0.5f * dt;
            eint[i] += DT * (deintdt[i] - deintdt_in[i]);
        }
    }
}

/* ---- cfd/Boundary/Inlet/Inlet.cl, cfd/Boundary/Outlet/Outlet.cl, cfd/Boundary/Portal/Mirror.cl: the
 * element-wise kernels of the open-boundary presets (cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml).
 * Vectors keep every stored component (w in 3-D), sums run left to right like the scripts write them. */

/* Inlet.cl:64-126.  inlet_N: (usize, usize). */
void aqo_inlet_feed(const aqo_defs* D, int* imove, const aqo_usize* iset, float* r, float* u, float* dudt,
                    float* rho, float* drhodt, float* m, float* p, const float* refd, aqo_usize N,
                    aqo_usize nbuffer, float cs, float p0, const float* g, float dr, const float* inlet_r,
                    const float* inlet_ru, const float* inlet_rv, const aqo_usize* inlet_N, const float* inlet_n,
                    float inlet_U, const float* inlet_rFS, float inlet_R, int inlet_starving)
{
    const int dims = D->dims, vs = VS(dims);
    if (inlet_starving == 0)
        return;
    const float off = inlet_R - D->SUPPORT * D->H - 0.5f * dr;
    aqo_usize n = nbuffer;
    if (inlet_N[0] * inlet_N[1] < n)
        n = inlet_N[0] * inlet_N[1];
    AQO_FOR_I(n) {
        const aqo_usize ii = N - nbuffer + i;
        float u_fac, v_fac;
        if (dims == 2) {
            u_fac = ((float)i + 0.5f) / (float)inlet_N[0];
            v_fac = 0.f;
        } else {
            const aqo_usize u_id = i % inlet_N[0], v_id = i / inlet_N[0];
            u_fac = ((float)u_id + 0.5f) / (float)inlet_N[0];
            v_fac = ((float)v_id + 0.5f) / (float)inlet_N[1];
        }
        float d[4];
        for (int k = 0; k < vs; k++) {
            const float rk = inlet_r[k] + u_fac * inlet_ru[k] + v_fac * inlet_rv[k] + off * inlet_n[k];
            r[(size_t)ii * vs + k] = rk;
            dudt[(size_t)ii * vs + k] = 0.f;
            u[(size_t)ii * vs + k] = inlet_U * inlet_n[k];
            d[k] = rk - inlet_rFS[k];
        }
        imove[ii] = 1;
        drhodt[ii] = 0.f;
        const float rd = refd[iset[ii]];
        const float ph = rd * aqo_dotv(g, d, vs);
        m[ii] = (dims == 3) ? rd * dr * dr * dr : rd * dr * dr;
        rho[ii] = rd + ph / (cs * cs);
        p[ii] = ph + p0;
    }
}

/* Inlet.cl:148-176 */
void aqo_inlet_rates(const int* imove, const float* r, float* u, float* dudt, float* drhodt, aqo_usize N,
                     const float* inlet_r, float inlet_U, const float* inlet_n, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = r[(size_t)i * vs + k] - inlet_r[k];
        if (aqo_dotv(d, inlet_n, vs) > 0.f)
            continue;
        for (int k = 0; k < vs; k++) {
            u[(size_t)i * vs + k] = inlet_U * inlet_n[k];
            dudt[(size_t)i * vs + k] = 0.f;
        }
        drhodt[i] = 0.f;
    }
}

/* Outlet.cl:53-96 */
void aqo_outlet_rates(const int* imove, const aqo_usize* iset, const float* r, float* u, float* rho, float* p,
                      float* dudt, float* dudt_in, float* drhodt, float* drhodt_in, const float* refd, aqo_usize N,
                      float cs, float p0, const float* g, const float* outlet_r, const float* outlet_n,
                      float outlet_U, const float* outlet_rFS, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        float d[4], e[4];
        for (int k = 0; k < vs; k++) {
            d[k] = r[(size_t)i * vs + k] - outlet_r[k];
            e[k] = r[(size_t)i * vs + k] - outlet_rFS[k];
        }
        if (aqo_dotv(d, outlet_n, vs) < 0.f)
            continue;
        drhodt[i] = 0.f;
        drhodt_in[i] = 0.f;
        for (int k = 0; k < vs; k++) {
            dudt[(size_t)i * vs + k] = 0.f;
            dudt_in[(size_t)i * vs + k] = 0.f;
            u[(size_t)i * vs + k] = outlet_U * outlet_n[k];
        }
        const float rd = refd[iset[i]];
        const float ph = rd * aqo_dotv(g, e, vs);
        rho[i] = rd + ph / (cs * cs);
        p[i] = ph + p0;
    }
}

/* Outlet.cl:108-131 */
void aqo_outlet_feed(const aqo_defs* D, int* imove, float* r_in, aqo_usize N, const float* domain_max,
                     const float* outlet_r, const float* outlet_n)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(N) {
        if (imove[i] != 1)
            continue;
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = r_in[(size_t)i * vs + k] - outlet_r[k];
        const float dist = aqo_dotv(d, outlet_n, vs);
        if (dist < 0.f)
            continue;
        if (dist > D->SUPPORT * D->H) {
            for (int k = 0; k < vs; k++) /* VEC_ONE: w = 0 in 3-D (types/3D.h:38) */
                r_in[(size_t)i * vs + k] = domain_max[k] + ((k < dims) ? 1.f : 0.f);
            imove[i] = -256;
        }
    }
}

/* Portal/Mirror.cl:32-50: the cell of a moved particle, LinkList.cl.in's formula */
static aqo_usize aqo_portal_cell(const aqo_defs* D, const float* r, const float* r_min, const aqo_usize* n_cells)
{
    const float idist = 1.f / (D->SUPPORT * D->H);
    const aqo_usize cx = (aqo_usize)((r[0] - r_min[0]) * idist) + 3u;
    const aqo_usize cy = (aqo_usize)((r[1] - r_min[1]) * idist) + 3u;
    if (D->dims == 3) {
        const aqo_usize cz = (aqo_usize)((r[2] - r_min[2]) * idist) + 3u;
        return cx - 1u + (cy - 1u) * n_cells[0] + (cz - 1u) * n_cells[0] * n_cells[1];
    }
    return cx - 1u + (cy - 1u) * n_cells[0];
}

/* Portal/Mirror.cl:70-95 */
void aqo_portal_mirror(const aqo_defs* D, float* r, int* imirrored, aqo_usize* icell, aqo_usize N,
                       const float* portal_in_r, const float* portal_out_r, const float* portal_n,
                       const float* r_min, const aqo_usize* n_cells)
{
    const int vs = VS(D->dims);
    AQO_FOR_I(N) {
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = r[(size_t)i * vs + k] - portal_out_r[k];
        if (fabsf(aqo_dotv(d, portal_n, vs)) > D->SUPPORT * D->H) {
            imirrored[i] = 0;
            continue;
        }
        imirrored[i] = 1;
        for (int k = 0; k < vs; k++)
            r[(size_t)i * vs + k] = portal_in_r[k] + d[k];
        icell[i] = aqo_portal_cell(D, r + (size_t)i * vs, r_min, n_cells);
    }
}

/* Portal/Mirror.cl:108-124 */
void aqo_portal_unmirror(float* r, const int* imirrored, aqo_usize N, const float* portal_in_r,
                         const float* portal_out_r, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        if (!imirrored[i])
            continue;
        for (int k = 0; k < vs; k++)
            r[(size_t)i * vs + k] = portal_out_r[k] + (r[(size_t)i * vs + k] - portal_in_r[k]);
    }
}

/* Portal/Mirror.cl:136-153 */
void aqo_portal_teleport(float* r, aqo_usize N, const float* portal_in_r, const float* portal_out_r,
                         const float* portal_n, int dims)
{
    const int vs = VS(dims);
    AQO_FOR_I(N) {
        float d[4];
        for (int k = 0; k < vs; k++)
            d[k] = r[(size_t)i * vs + k] - portal_out_r[k];
        if (aqo_dotv(d, portal_n, vs) < 0.f)
            continue;
        for (int k = 0; k < vs; k++)
            r[(size_t)i * vs + k] = portal_in_r[k] + d[k];
    }
}

/* cfd/Boundary/Portal/Shepard.cl:44-113: the particles mirrored to the in portal (imirrored) add the fluid
 * they see there to their Shepard factor.  LOCAL_MEM_SIZE build: starts from shepard[i], stored at the end. */
void aqo_portal_shepard(const aqo_defs* D, const aqo_ll* L, const int* imove, const int* imirrored,
                        const float* r, const float* rho, const float* m, float* shepard)
{
    const int dims = D->dims, vs = VS(dims);
    AQO_FOR_I(L->N) {
        if ((imove[i] < -3) || (imove[i] > 1) || (!imirrored[i]))
            continue;
        const float* r_i = r + (size_t)i * vs;
        float s = shepard[i];
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((imove[j] != 1) || (imirrored[j]))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            s += kernelW(q, dims) * D->CONW * m[j] / rho[j];
        }
        NEIGHS_END
        shepard[i] = s;
    }
}

/* cfd/Boundary/Portal/Interactions.cl:47-146 (morris: the __LAP_MORRIS__ branch, :128-129): the mirrored
 * fluid particles add the interactions with what they see at the in portal to grad_p / lap_u / div_u. */
void aqo_portal_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove, const int* imirrored,
                             const float* r, const float* u, const float* rho, const float* m, const float* p,
                             float* grad_p, float* lap_u, float* div_u, int morris)
{
    const int dims = D->dims, vs = VS(dims);
    const float cleary = (dims == 3) ? 10.f : 8.f;
    const float H = D->H;
    AQO_FOR_I(L->N) {
        if ((!imirrored[i]) || (imove[i] != 1))
            continue;
        const float* r_i = r + (size_t)i * vs;
        const float* u_i = u + (size_t)i * vs;
        const float p_i = p[i], rho_i = rho[i];
        float gp[3], lu[3], du = div_u[i];
        for (int d = 0; d < dims; d++) {
            gp[d] = grad_p[(size_t)i * vs + d];
            lu[d] = lap_u[(size_t)i * vs + d];
        }
        NEIGHS_BEGIN(L, i, dims)
        {
            if ((imirrored[j]) || ((imove[j] != 1) && (imove[j] != -1)))
                continue;
            float r_ij[3], q;
            if (!pair_q(D, r_i, r + (size_t)j * vs, r_ij, &q))
                continue;
            const float rho_j = rho[j], m_j = m[j], p_j = p[j];
            float u_ij[3];
            for (int d = 0; d < dims; d++)
                u_ij[d] = u[(size_t)j * vs + d] - u_i[d];
            const float udr = dotv(u_ij, r_ij, dims);
            const float f_ij = kernelF(q, dims) * D->CONF * m_j;
            const float pf = (p_i + p_j) / (rho_i * rho_j) * f_ij;
            if (morris) {
                const float lf = f_ij * 2.f / (rho_i * rho_j);
                for (int d = 0; d < dims; d++) {
                    gp[d] += pf * r_ij[d];
                    lu[d] += lf * u_ij[d];
                }
            } else {
                const float r2 = (q * q + 0.01f) * H * H;
                const float lf = f_ij * cleary * udr / (r2 * rho_i * rho_j);
                for (int d = 0; d < dims; d++) {
                    gp[d] += pf * r_ij[d];
                    lu[d] += lf * r_ij[d];
                }
            }
            du += udr * f_ij * rho_i / rho_j;
        }
        NEIGHS_END
        for (int d = 0; d < dims; d++) {
            grad_p[(size_t)i * vs + d] = gp[d];
            lap_u[(size_t)i * vs + d] = lu[d];
        }
        div_u[i] = du;
    }
}
