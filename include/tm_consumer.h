/*
 * include/tm_consumer.h -- C ABI of two consumer-side operators of the SPB = 2 (vegetation + grazer) Rietkerk model on B200
 * (sm_100a): the first part of SURVEY 8 f3.  Plain pointers and sizes only; no CPU fallback (tm_cons_create fails with
 * TM_E_CUDA without a device).  R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.
 *
 *   tm_cons_gamma_host   replaces calc_gamma_2Sp_NonRefugia(centralNeighboringSites, Rho_t, gamma, Rho_avg, rfrac,
 *                        nVeg_frac, r_max, L)   R/:2530-2603 (rietkerk_bjork_basic.h declares the 3-species twin at :371)
 *   tm_cons_rk4_host     replaces RK4_Integrate_Stochastic_2Sp(Rho_t, Rho_tsar, K1..K4, nR2, a, c, gmax, alpha, rW, W0,
 *                        Dxd2, K, A, H, E, t, dt, dx, g)   R/:2472-2527 with f_2Dor_2Sp R/:2445-2470
 *   tm_cons_grazer_dornic_host   replaces the grazer's Dornic loop with upwind advection inlined in
 *                        rietkerk_Dornic_2D_MultiSp, R/:3908-4114 (s = 1)
 *   tm_cons_step2_*      the whole SPB = 2 time step resident on the device (below)
 * Not yet behind this ABI (remaining): the predator (SPB = 3) variants.
 */
#ifndef TM_CONSUMER_H
#define TM_CONSUMER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field order: [0] = P (vegetation), [1] = G (grazer), [2] = W (soil water), [3] = O (surface water): the reference's
 * SPB = 2 memory order.  Dw, Do are D[2]/dx^2 and D[3]/dx^2 (Dxd2[s], R/:3463-3480). */
typedef struct tm_cons_params {
    int32_t L, pad_;
    double dt, a, c, gmax, alpha, rW, W0, K1, K2, Dw, Do;
    double A01, H01, E1;       /* A[0][1], H[0][1], E[1] */
} tm_cons_params;

typedef struct tm_cons_ctx tm_cons_ctx;

int tm_cons_create(int L, int device, tm_cons_ctx **out);
int tm_cons_destroy(tm_cons_ctx *);
const char *tm_cons_last_error(const tm_cons_ctx *);   /* ctx == NULL: error of the last failed tm_cons_create on this thread */

/* nsteps RK4 steps in place on fields[4][L*L] (host buffer: uploaded, stepped on the device, downloaded). */
int tm_cons_rk4_host(tm_cons_ctx *, const tm_cons_params *p, double *fields, int nsteps);

/* The grazer's perception field.  prey: [L*L] vegetation; offsets: [n_offsets][2] = {row offset, column offset}, the first
 * n_offsets entries of the distance-sorted disc list that are summed (the reference's eff_nR_Perp_size, R/:2569);
 * rho_avg_prey / rho_avg_self: spatial means of vegetation and grazer; gamma: [L*L] in/out -- the disc sum is added to the
 * value the array holds on entry, as the reference does (R/:2581). */
int tm_cons_gamma_host(tm_cons_ctx *, const double *prey, const int32_t *offsets, int n_offsets, double rho_avg_prey,
                       double rho_avg_self, double *gamma);

/* The grazer's Dornic update (rho* of species 1).  D, v, M, sigma: D[1], v[1], M[1], sigma[1]; D0_floor = D[0]/dx^2
 * (diff_coefficient[0], the floor of the advection-corrected diffusion, R/:4014). */
typedef struct tm_graz_params {
    int32_t L, pad_;
    double dt, dx, D, v, M, sigma, D0_floor;
} tm_graz_params;

/* prev: DRho[1] [L*L]; gamma: [L*L] in/out (clamped at 1 as R/:3915 does); rho_star: [L*L] out; v_eff: [L*L][2] out (the
 * advection vector drawn for every cell).  Random numbers: replay_u == NULL: Philox4x32-10 keyed by (seed, trial, step,
 * cell) -- uniform 0 of a cell's stream is its advection direction, the sampler draws follow; otherwise cell i takes its k-th
 * uniform from replay_u[i*stride + k] (draws past `stride` are counted in *overflow and return 0.75).  The caller applies
 * the reference's whole-species skip `if (Rhox_avg[s] == 0.0) continue;` (R/:3910). */
int tm_cons_grazer_dornic_host(tm_cons_ctx *, const tm_graz_params *p, uint64_t seed, uint32_t trial, uint32_t step,
                               const double *replay_u, uint32_t stride, const double *prev, double *gamma, double *rho_star,
                               double *v_eff, uint64_t *overflow);

/* ---- The SPB = 2 time step, resident on the device ------------------------------------------------------------------
 * Replaces the body of the time loop of rietkerk_Dornic_2D_MultiSp compiled with -DSPB=2, R/:3752-4183: vegetation Dornic
 * (R/:3752-3809) -> spatial means and vegetation cover of DRho (R/:3813-3824) -> calc_gamma_2Sp_NonRefugia (R/:3844) ->
 * grazer Dornic with upwind advection (R/:3908-4114; skipped when <G> == 0, R/:3910) -> RK4_Integrate_Stochastic_2Sp
 * (R/:4137) -> clamp and commit (R/:4150-4181).  Six launches per step on the context's stream and no host round trip: the
 * scalars the reference derives between its loops (<P>, <G>, cover -> disc size) are reduced on the device and read there.
 * The means come from a fixed summation tree, not the reference's left-to-right sum: reproducible, a few ulp from the
 * sequential value (gamma inherits that: <= 1e-13 relative).
 *
 * setup: veg = the vegetation's constants as the SPB = 1 step takes them (struct tm_params of tm_lattice.h: L, dt, dx, d,
 *   D0, sigma0, clamp_eps, clamp_lo are read); pg, pc as above; central = the full distance-sorted disc list
 *   computeNeighboringSitesCentral(int(r_max)) as [n_central][2] = {row offset, column offset}; fr_graz = pR[1] /
 *   (dx r_max_effective) (R/:3409).  upload: fields [4][L*L] = P, G, W, O; gamma [L*L] or NULL (zeros; the reference never
 *   resets the array between steps, so it is state).
 * run: nsteps steps with Philox streams keyed (seed, trial, step0 + s, cell); the grazer's streams use seed ^
 *   0x9E3779B97F4A7C15.  Asynchronous: returns after enqueueing.  replay: ONE step with host-supplied uniforms, cell i of
 *   the vegetation / the grazer taking its k-th from u_veg[i*stride + k] / u_graz[i*stride + k].
 * download (synchronises): any of fields [4][L*L], gamma [L*L], v_eff [L*L][2] (the last step's advection vectors),
 *   means[4] = {<P>, <G>, #{P > 0}, disc offsets summed} of the state the LAST step started from, overflow (replay draws
 *   past `stride`, cumulative since upload) may be NULL. */
struct tm_params;
int tm_cons_step2_setup(tm_cons_ctx *, const struct tm_params *veg, const tm_graz_params *pg, const tm_cons_params *pc,
                        const int32_t *central, int n_central, double fr_graz);
int tm_cons_step2_upload(tm_cons_ctx *, const double *fields, const double *gamma);
/* RK4 part of the step: 0 (default) = one launch for the four stages and the clamp (tile-fused, six launches per step);
 * 1 = one launch per stage (nine per step).  Bit-identical results; the second form is kept as the cross-check. */
int tm_cons_step2_set_rk4_stages(tm_cons_ctx *, int on);
/* Means of the step: 0 (default) = the device reduction tree (reproducible, a few ulp from the reference's left-to-right sums, so
 * that the step equals the oracle bit for bit GIVEN those means); 1 = strict: mean_of_vector's own left-to-right sums (R/:1211-1219) by one
 * thread per field -- the step then equals the reference-pinned oracle bit for bit free-running, at the cost of a dependent chain of L*L
 * additions per step (parity runs; ~70 ms at L = 4096). */
int tm_cons_step2_set_strict_means(tm_cons_ctx *, int on);
/* Rietkerk_FPE flavour of the resident step: the body of the time loop of rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=2,
 * Rietkerk_FPE/rietkerk_bjork_FKE.cpp:4094-4462 -- perception field with calc_gamma_vel_2Sp_NonRefugia's rules (:2534-2607: a
 * vanished consumer gets gamma = 1) -> vegetation Dornic (:4154-4211) -> four-field RK4 over f_DorFKE_2Sp (:2449-2531: the grazer's
 * mortality M1 sits in the reaction terms) -> advdiff_FKE_MultiSp on the grazer (:4308; the gather-form kernels of tm_fke.h, on the
 * device-resident field: the reference's per-step cudaMemcpy pair :4282-4299 has no counterpart) -> clamp `v < eps && v >= lo`
 * with the veg->clamp_eps / clamp_lo given to setup (1e-10 and 0 in the reference, :4458).  The grazer has no Dornic update and
 * the advection vectors keep what the context holds (zeros after setup: the two-species runner never assigns them).
 * t: the grazer's movement windows and stencil sizes (struct tm_fke_terms of tm_fke.h, CPU-path convention F/:3640-3641);
 * gauss_D / gauss_V: its two stencils (host).  Ten launches per step.  t == NULL switches back to the Rietkerk_FD flavour. */
struct tm_fke_terms;
int tm_cons_step2_set_fpe(tm_cons_ctx *, const struct tm_fke_terms *t, const double *gauss_D, const double *gauss_V, double M1);
int tm_cons_step2_run(tm_cons_ctx *, uint64_t seed, uint32_t trial, uint32_t step0, int nsteps);
int tm_cons_step2_replay(tm_cons_ctx *, const double *u_veg, const double *u_graz, uint32_t stride);
int tm_cons_step2_download(tm_cons_ctx *, double *fields, double *gamma, double *v_eff, double *means, uint64_t *overflow);

/* Device time (ms) of the kernels of the last call and how many launches that was. */
int tm_cons_last_ms(tm_cons_ctx *, float *ms, int *n_launches);

#ifdef __cplusplus
}
#endif
#endif
