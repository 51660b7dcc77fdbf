/* aqo.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the AQUAgpusph 5.0.4 per-time-step particle pipeline,
 * used only as the checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product path
 * (aquagpusph_b200/csrc -> libaquacuda.so) never links, imports or calls
 * anything in this directory.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Conventions shared by every function:
 *   - dims is 2 or 3; a "vec" array has stride vs = (dims == 3) ? 4 : 2 floats
 *     (resources/Scripts/types/2D.h:23-31, 3D.h:23-31), w unused in 3D.
 *   - a "matrix" array has stride ms = (dims == 3) ? 16 : 4 floats.
 *   - indices are 32-bit unsigned (the reference default, State.cpp:499-502).
 *   - fp32 arithmetic, compiled with -ffp-contract=off (no FMA contraction).
 *
 * Parity pin: aqo_linklist is checked against the property tests the
 * reference ships (tests/{2D,3D}/{LinkList,RadixSort}/cMake/check.py) on the
 * reference's own particles.dat inputs, and every physics kernel is checked
 * against the reference's unmodified .cl sources compiled as C++ behind a shim
 * (oracle/ref_shim -> oracle/_ref/).  See DESIGN.md "Oracle".
 */
#ifndef AQO_H
#define AQO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint32_t aqo_usize;

/* Compile-time "-D" definitions the reference bakes into every kernel
 * (resources/Presets/src/basic.xml:119-123, CalcServer.cpp:245-257).  H, CONW
 * and CONF are the 6-significant-digit roundings of h, 1/h^d, 1/h^(d+2). */
typedef struct {
    int dims;      /* 2 or 3 */
    float H;       /* -DH */
    float CONW;    /* -DCONW */
    float CONF;    /* -DCONF */
    float SUPPORT; /* -DSUPPORT=2.f */
} aqo_defs;

/* printf("%#G") + "f" rounding of an evaluated <Define> (CalcServer.cpp:245-257) */
float aqo_define_round6(float v);
/* Fill defs from the full precision h as basic.xml:119-123 does. */
void aqo_make_defs(aqo_defs* d, int dims, float h);

/* Restrict the particle loops of the kernels in aqo_kernels.c to rows [lo, hi) for
 * the CALLING thread (thread-local; default: every row).  Lets a driver spread one
 * kernel over several host threads for the CPU baseline. */
void aqo_set_range(aqo_usize lo, aqo_usize hi);

/* ---- link-list (aquagpusph/CalcServer/LinkList.cpp:326-494) ------------- */
void aqo_minmax(const float* r, aqo_usize N, int dims, float* rmin, float* rmax);
int aqo_ncells(const float* rmin, const float* rmax, int dims, float support,
               float h, aqo_usize ncells[4]);
void aqo_icell(aqo_usize* icell, const float* r, aqo_usize N, int dims,
               const float* rmin, float support, float h,
               const aqo_usize ncells[4]);
void aqo_radix_sort(aqo_usize* keys, aqo_usize n, aqo_usize* perm,
                    aqo_usize* inv_perm);
void aqo_ihoc(const aqo_usize* icell_sorted, aqo_usize N, aqo_usize* ihoc,
              aqo_usize n_cells_w);
/* whole LinkList::_execute; ihoc must hold >= n_cells.w entries, checked
 * against ihoc_capacity (returns -1 when it does not fit, leaving ncells set) */
int aqo_linklist(const float* r, aqo_usize N, int dims, float support, float h,
                 int recompute_grid, float* rmin, float* rmax,
                 aqo_usize ncells[4], aqo_usize* icell, aqo_usize* ihoc,
                 size_t ihoc_capacity, aqo_usize* perm, aqo_usize* inv_perm);
/* generic permutation out[idx[i]] = in[i]  (basic/Sort.cl:57-124, UnSort.cl.in:30-42) */
void aqo_scatter(void* out, const void* in, const aqo_usize* idx, aqo_usize N,
                 size_t elem_bytes);

/* ---- element-wise kernels ------------------------------------------------ */
void aqo_eos(const aqo_usize* iset, const int* imove, const float* rho,
             float* p, const float* refd, aqo_usize N, float cs, float p0);
void aqo_rates(const aqo_usize* iset, const int* imove, const float* rho,
               const float* grad_p, const float* lap_u, const float* div_u,
               float* dudt, float* drhodt, const float* visc_dyn, aqo_usize N,
               const float* g, int dims);
void aqo_timestep(const int* imove, const float* u, float* dt_var, aqo_usize N,
                  float dt, float dt_min, float courant, float dt_Ma, float h,
                  int dims);
float aqo_reduce_min(const float* v, aqo_usize N);
void aqo_domain(int* imove, float* r_in, float* u_in, float* dudt_in, float* m,
                aqo_usize N, const float* domain_min, const float* domain_max,
                int dims);
void aqo_binormal(const float* normal, float* tangent, float* binormal,
                  aqo_usize N, int dims);

/* time schemes (resources/Scripts/basic/time_scheme/) */
void aqo_euler_predictor(const float* r, const float* u, const float* dudt,
                         const float* rho, const float* drhodt, float* r_in,
                         float* u_in, float* dudt_in, float* rho_in,
                         float* drhodt_in, aqo_usize N, int dims);
void aqo_euler_corrector(const int* imove, float* r, float* u,
                         const float* dudt, float* rho, const float* drhodt,
                         aqo_usize N, float dt, int dims);
void aqo_ie_predictor(const int* imove, const float* r, const float* u,
                      const float* dudt, const float* rho, const float* drhodt,
                      float* r_in, float* u_in, float* dudt_in, float* rho_in,
                      float* drhodt_in, aqo_usize N, float dt, int dims);
void aqo_ie_corrector(const int* imove, float* r, float* u, const float* dudt,
                      float* rho, const float* drhodt, const float* dudt_in,
                      const float* drhodt_in, aqo_usize N, float dt, int dims);
void aqo_mp_predictor(const float* r, const float* u, const float* dudt,
                      const float* rho, const float* drhodt, float* r_in,
                      float* u_in, float* dudt_in, float* rho_in,
                      float* drhodt_in, aqo_usize N, int dims);
void aqo_mp_midpoint(const int* imove, const float* u_in, float* u,
                     const float* dudt, const float* rho_in, float* rho,
                     const float* drhodt, aqo_usize N, float dt, int dims);
void aqo_mp_relax(const int* imove, const float* dudt_in, float* dudt,
                  const float* drhodt_in, float* drhodt, aqo_usize N,
                  float relax, int dims);
void aqo_mp_residuals(const int* imove, const float* m, const float* u,
                      const float* dudt_in, const float* dudt, const float* rho,
                      const float* p, const float* drhodt_in,
                      const float* drhodt, float* residual, aqo_usize N,
                      int dims);
void aqo_mp_corrector(const int* imove, const float* r_in, float* r,
                      const float* u_in, float* u, const float* dudt,
                      const float* rho_in, float* rho, const float* drhodt,
                      aqo_usize N, float dt, int dims);

/* ---- neighbour sweeps ---------------------------------------------------- */
typedef struct {
    const aqo_usize* icell;
    const aqo_usize* ihoc;
    aqo_usize ncells[4];
    aqo_usize N;
} aqo_ll;

void aqo_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove,
                      const float* r, const float* u, const float* rho,
                      const float* m, const float* p, float* grad_p,
                      float* lap_u, float* div_u);
/* mode 0: basic/Shepard.cl (EXCLUDED = imove>=3); 1: cfd/Shepard.cl (imove!=1) */
void aqo_shepard(const aqo_defs* D, const aqo_ll* L, int cfd_mode,
                 const int* imove, const float* r, const float* rho,
                 const float* m, float* shepard);
void aqo_neighs(const aqo_ll* L, const int* imove, aqo_usize* n_neighs,
                aqo_usize neighs_limit, int dims);
void aqo_sensors(const aqo_defs* D, const aqo_ll* L, const int* imove,
                 const float* r, const float* m, float* u, float* rho,
                 float* p);
void aqo_sensors_renorm(const int* imove, const float* shepard, float* u,
                        float* rho, float* p, aqo_usize N, int dims);

/* delta-SPH (basic/deltaSPH.cl via cfd/deltaSPH.cl: EXCLUDED = imove != 1) */
void aqo_dsph_simple(const aqo_usize* iset, const int* imove, float* lap_p_corr,
                     const float* refd, aqo_usize N, const float* g, int dims);
void aqo_dsph_full(const aqo_defs* D, const aqo_ll* L, const int* imove,
                   const float* r, const float* rho, const float* m,
                   const float* p, float* lap_p_corr);
void aqo_dsph_full_mls(const int* imove, const float* mls, float* lap_p_corr,
                       aqo_usize N, int dims);
void aqo_dsph_lapp(const aqo_defs* D, const aqo_ll* L, const int* imove,
                   const float* r, const float* rho, const float* m,
                   const float* p, float* lap_p);
void aqo_dsph_lapp_corr(const aqo_defs* D, const aqo_ll* L, const int* imove,
                        const float* r, const float* rho, const float* m,
                        const float* lap_p_corr, float* lap_p);
void aqo_dsph_apply(const aqo_usize* iset, const int* imove, const float* rho,
                    const float* lap_p, float* drhodt, const float* refd,
                    const float* delta, aqo_usize N, float dt);

/* MLS (basic/MLS.cl) */
void aqo_mls(const aqo_defs* D, const aqo_ll* L, const int* imove,
             const float* r, const float* rho, const float* m, float* mls,
             aqo_usize mls_imove);
void aqo_mls_inv(const int* imove, float* mls, aqo_usize N,
                 aqo_usize mls_imove, int dims);

/* BIe boundary integrals (cfd/Boundary/BIe/) */
void aqo_bie_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove,
                          const float* r, const float* normal, const float* u,
                          const float* m, float* grad_w_bi, float* div_u_bi);
void aqo_bie_p_boundary(const aqo_defs* D, const aqo_ll* L, const int* imove,
                        const float* r, const float* m, const float* rho,
                        float* p);
void aqo_bie_rates(const int* imove, const float* rho, const float* p,
                   const float* u, const float* grad_w_bi,
                   const float* div_u_bi, float* grad_p, float* div_u,
                   aqo_usize N, int dims);
void aqo_bie_filter_press(const aqo_usize* iset, const int* imove, float* p,
                          aqo_usize forces_iset, aqo_usize N);
void aqo_bie_force_press(const int* imove, const float* r, const float* normal,
                         const float* m, const float* p, float* force_p,
                         float* moment_p, const float* forces_r, aqo_usize N,
                         int dims);
void aqo_bie_elastic_bounce(const aqo_ll* L, const int* imove,
                            const float* r_in, const float* normal,
                            const float* m, const float* u_in, float* dudt,
                            float dt, float dr_factor, int dims);
void aqo_bie_force_bound(const int* imove, const float* m,
                         const float* dudt_preelastic,
                         const float* dudt_elastic, float* force_elastic,
                         aqo_usize N, int dims);
void aqo_bie_pst(const aqo_ll* L, const int* imove, float* r,
                 const float* normal, const float* m, const float* rho,
                 float DIMS_define, float dr_factor, int dims);

/* Reduction tool, sum in the reference's tree order (Reduction.cl.in:35-66,
 * Reduction.cpp:376-436) with work-group size wg (power of two). */
float aqo_reduce_sum_tree(const float* v, aqo_usize N, aqo_usize wg);
void aqo_reduce_sum_vec_tree(const float* v, aqo_usize N, int ncomp,
                             aqo_usize wg, float* out);
float aqo_reduce_max(const float* v, aqo_usize N);
aqo_usize aqo_reduce_max_u32(const aqo_usize* v, aqo_usize N);

/* cfd/Motions/{Transform,UnTransform,Velocity,Acceleration}.cl (preset cfd/motion.xml) */
void aqo_motion_transform(const unsigned* iset, const int* imove, float* r, float* normal,
                          float* tangent, aqo_usize N, unsigned motion_iset,
                          const float* motion_r, const float* motion_a, int dims);
void aqo_motion_untransform(const unsigned* iset, const int* imove, float* r, float* normal,
                            float* tangent, aqo_usize N, unsigned motion_iset,
                            const float* motion_r_in, const float* motion_a_in, int dims);
void aqo_motion_velocity(const unsigned* iset, const int* imove, const float* r, float* u,
                         aqo_usize N, unsigned motion_iset, const float* motion_drdt,
                         const float* motion_a, const float* motion_dadt, int dims);
void aqo_motion_acceleration(const unsigned* iset, const int* imove, const float* r, float* dudt,
                             aqo_usize N, unsigned motion_iset, const float* motion_ddrddt,
                             const float* motion_a, const float* motion_ddaddt, int dims);

/* cfd/Energy/Energy.cl::power, ::energy (preset cfd/energy.xml) */
void aqo_energy_power(float* energy_dekdt, float* energy_depdt, float* energy_decdt, const int* imove,
                      const float* u, const float* rho, const float* m, const float* p,
                      const float* dudt, const float* drhodt, aqo_usize N, const float* g, int dims);
void aqo_energy_energy(float* energy_ek, float* energy_ep, float* energy_ec, const unsigned* iset,
                       const int* imove, const float* r, const float* u, const float* rho,
                       const float* m, const float* refd, aqo_usize N, const float* g, float cs,
                       int dims);

/* cfd/Energy/EnergyKin.cl:38-54 (preset cfd/energy_kin.xml) */
void aqo_energy_kin(float* energy_kin, const int* imove, const float* u, const float* m, aqo_usize N, int dims);
/* cfd/Forces/Forces.cl:50-84 (preset cfd/forces.xml); forces_m is vec4 in 2-D and 3-D */
void aqo_forces(float* forces_f, float* forces_m, const int* imove, const float* r, const float* dudt,
                const float* m, aqo_usize N, const float* g, const float* forces_r, int dims);
/* basic/DensityClamp.cl:41-52 (preset basic/densityClamp.xml) */
void aqo_density_clamp(float* rho_in, aqo_usize N, float rho_min, float rho_max);
/* basic/IdInverse.cl:33-42 (preset basic/id_inverse.xml) */
void aqo_id_inverse(const aqo_usize* id, aqo_usize* id_inverse, aqo_usize N);

/* basic/time_scheme/adam_bashforth.cl (preset basic/time_scheme/adams_bashforth.xml): ::sort :122-146,
 * ::corrector :207-255, ::postcorrector :281-309; ::predictor :93-113 is aqo_mp_predictor's body.
 * steps = TSCHEME_ADAMS_BASHFORTH_STEPS (5 when undefined, :66-68) */
void aqo_ab_sort(const float* const* dudt_as_in, float* const* dudt_as, const float* const* drhodt_as_in,
                 float* const* drhodt_as, const aqo_usize* id_sorted, aqo_usize N, int dims);
void aqo_ab_corrector(const int* imove, float* r, float* u, const float* dudt, float* rho, const float* drhodt,
                      const float* const* dudt_as, const float* const* drhodt_as, aqo_usize N, float dt,
                      unsigned iter, unsigned steps, int dims);
void aqo_ab_postcorrector(const float* const* dudt_as, const float* const* drhodt_as, const float* dudt,
                          const float* drhodt, float* const* dudt_as_in, float* const* drhodt_as_in,
                          aqo_usize N, int dims);

/* cfd/Boundary/BI/NoSlip.cl:52-130 (preset cfd/BINoSlip.xml), __LAP_MONAGHAN__ */
void aqo_bi_noslip(const aqo_defs* D, const aqo_ll* L, const unsigned* iset, const int* imove, const float* r,
                   const float* normal, const float* u, const float* rho, const float* m, float* lap_u,
                   unsigned noslip_iset, float dr);

#ifdef __cplusplus
}
#endif
/* Remote (halo) terms of MLS and delta-SPH (ours: see aqo_kernels.c) */
void aqo_mpi_mls(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i, const int* imove,
                 const float* r, const float* mpi_r, const float* mpi_rho, const float* mpi_m,
                 float* mls, aqo_usize mls_imove);
void aqo_mpi_dsph_full_lapp(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i,
                            const int* imove, const float* r, const float* p, const float* mpi_r,
                            const float* mpi_rho, const float* mpi_m, const float* mpi_p,
                            float* lap_p_corr, float* lap_p);
void aqo_mpi_dsph_lapp_corr(const aqo_defs* D, const aqo_ll* L, const aqo_usize* icell_i,
                            const int* imove, const float* r, const float* lap_p_corr,
                            const float* mpi_r, const float* mpi_rho, const float* mpi_m,
                            const float* mpi_lap_p_corr, float* lap_p);

/* cfd/Boundary/Symmetry/Mirror.cl:37-251 (preset cfd/symmetry.xml) */
void aqo_sym_drop(int* imove, float* r, aqo_usize N, const float* symmetry_r, const float* symmetry_n,
                  const float* domain_max, int dims);
void aqo_sym_detect(const aqo_defs* D, const int* imove, const float* r_in, aqo_usize* imirror, aqo_usize N,
                    const float* symmetry_r, const float* symmetry_n);
void aqo_sym_feed(int* imove, int* iset, const aqo_usize* imirror, const aqo_usize* imirror_invperm,
                  aqo_usize* mirror_src, float* normal, float* tangent, float* r_in, aqo_usize N,
                  aqo_usize nbuffer, const float* symmetry_r, const float* symmetry_n, int dims);
void aqo_sym_set(const aqo_usize* mirror_src, float* m, float* u_in, float* dudt_in, float* dudt, float* rho_in,
                 float* drhodt_in, float* drhodt, aqo_usize N, const float* symmetry_n, int dims);
void aqo_sym_sort(const aqo_usize* mirror_src_in, aqo_usize* mirror_src, const aqo_usize* id_sorted, aqo_usize N);

/* cfd/ideal_gas: the element-wise kernels (EOS.cl:56-70, Rates.cl:53-68, Sort.cl:43-58, TimeStep.cl:62-97,
 * riemann/Rates.cl:39-53, time_scheme/midpoint.cl:47-144) */
void aqo_ig_eos(const aqo_usize* iset, const int* imove, const float* rho, const float* eint, float* p,
                const float* gamma, aqo_usize N);
void aqo_ig_rates(const int* imove, const float* rho, const float* p, const float* div_u, float* deintdt,
                  aqo_usize N);
void aqo_ig_sort(const float* eint_in, float* eint, const float* deintdt, float* deintdt_in,
                 const aqo_usize* id_sorted, aqo_usize N);
void aqo_ig_timestep(const aqo_defs* D, float* dt_var, const int* imove, const aqo_usize* iset, const float* u,
                     const float* rho, const float* p, aqo_usize N, float dt, float dt_min, float courant,
                     const float* div_u, const float* grad_p, const float* gamma);
void aqo_ig_riemann_rates(const int* imove, const float* work_density, float* deintdt, aqo_usize N);
void aqo_ig_mp_predictor(const float* eint, const float* deintdt, float* eint_in, float* deintdt_in, aqo_usize N);
void aqo_ig_mp_midpoint(const int* imove, const float* eint_in, const float* deintdt, float* eint, aqo_usize N,
                        float dt);
void aqo_ig_mp_relax(const int* imove, const float* deintdt_in, float* deintdt, aqo_usize N, float relax_midpoint);
void aqo_ig_mp_corrector(const int* imove, const float* eint_in, const float* deintdt, float* eint, aqo_usize N,
                         float dt);
/* cfd/ideal_gas/time_scheme/euler.cl:70-83, improved_euler.cl:49-97 (euler.cl's predictor = aqo_ig_mp_predictor) */
void aqo_ig_euler_corrector(const int* imove, float* eint, const float* deintdt, aqo_usize N, float dt);
void aqo_ig_ie_predictor(const int* imove, const float* eint, const float* deintdt, float* eint_in,
                         float* deintdt_in, aqo_usize N, float dt);
void aqo_ig_ie_corrector(const int* imove, const float* deintdt, const float* deintdt_in, float* eint,
                         aqo_usize N, float dt);
/* cfd/ideal_gas/riemann/Interactions.cl:50-168 */
void aqo_ig_riemann_interactions(const aqo_defs* D, const aqo_ll* L, const aqo_usize* iset, const int* imove,
                                 const float* r, const float* u, const float* rho, const float* m, const float* p,
                                 float* grad_p, float* div_u, float* work_density, const float* gamma);
/* cfd/ideal_gas/symmetry/Mirror.cl:32-48 */
void aqo_ig_sym_set(const aqo_usize* mirror_src, float* eint_in, float* deintdt_in, float* deintdt, aqo_usize N);

/* cfd/Boundary/Inlet/Inlet.cl:64-176, cfd/Boundary/Outlet/Outlet.cl:53-131, cfd/Boundary/Portal/Mirror.cl:70-153 */
void aqo_inlet_feed(const aqo_defs* D, int* imove, const aqo_usize* iset, float* r, float* u, float* dudt,
                    float* rho, float* drhodt, float* m, float* p, const float* refd, aqo_usize N,
                    aqo_usize nbuffer, float cs, float p0, const float* g, float dr, const float* inlet_r,
                    const float* inlet_ru, const float* inlet_rv, const aqo_usize* inlet_N, const float* inlet_n,
                    float inlet_U, const float* inlet_rFS, float inlet_R, int inlet_starving);
void aqo_inlet_rates(const int* imove, const float* r, float* u, float* dudt, float* drhodt, aqo_usize N,
                     const float* inlet_r, float inlet_U, const float* inlet_n, int dims);
void aqo_outlet_rates(const int* imove, const aqo_usize* iset, const float* r, float* u, float* rho, float* p,
                      float* dudt, float* dudt_in, float* drhodt, float* drhodt_in, const float* refd, aqo_usize N,
                      float cs, float p0, const float* g, const float* outlet_r, const float* outlet_n,
                      float outlet_U, const float* outlet_rFS, int dims);
void aqo_outlet_feed(const aqo_defs* D, int* imove, float* r_in, aqo_usize N, const float* domain_max,
                     const float* outlet_r, const float* outlet_n);
void aqo_portal_mirror(const aqo_defs* D, float* r, int* imirrored, aqo_usize* icell, aqo_usize N,
                       const float* portal_in_r, const float* portal_out_r, const float* portal_n,
                       const float* r_min, const aqo_usize* n_cells);
void aqo_portal_unmirror(float* r, const int* imirrored, aqo_usize N, const float* portal_in_r,
                         const float* portal_out_r, int dims);
void aqo_portal_teleport(float* r, aqo_usize N, const float* portal_in_r, const float* portal_out_r,
                         const float* portal_n, int dims);

/* cfd/Interactions.cl:60-145 under __LAP_FORMULATION__ == __LAP_MORRIS__ (:130-131) */
void aqo_interactions_morris(const aqo_defs* D, const aqo_ll* L, const int* imove, const float* r, const float* u,
                             const float* rho, const float* m, const float* p, float* grad_p, float* lap_u,
                             float* div_u);

/* cfd/Boundary/Portal/Shepard.cl:44-113, Portal/Interactions.cl:47-146 (morris: the __LAP_MORRIS__ branch) */
void aqo_portal_shepard(const aqo_defs* D, const aqo_ll* L, const int* imove, const int* imirrored,
                        const float* r, const float* rho, const float* m, float* shepard);
void aqo_portal_interactions(const aqo_defs* D, const aqo_ll* L, const int* imove, const int* imirrored,
                             const float* r, const float* u, const float* rho, const float* m, const float* p,
                             float* grad_p, float* lap_u, float* div_u, int morris);

#endif
