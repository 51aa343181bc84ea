/*
 * include/tm_consumer3.h -- C ABI of the SPB = 3 (vegetation + grazer + predator) time step on B200 (sm_100a), resident on the device
 * (SURVEY 8 f3).  Plain pointers and sizes only; no CPU fallback.  R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.
 *
 * Replaces the body of the time loop of rietkerk_Dornic_2D_MultiSp compiled with -DSPB=3, R/:3752-4183: vegetation Dornic
 * (R/:3752-3809) -> spatial means of vegetation, grazer, predator and the vegetation cover (R/:3813-3824) ->
 * calc_gamma_3Sp_NonRefugia (R/:3153-3268, call R/:3850) -> the consumers' Dornic loop with upwind advection for the grazer and the
 * predator (R/:3908-4114; a species whose mean is 0 is skipped, R/:3910) -> RK4_Integrate_Stochastic_MultiSp over f_2Dor_3Sp
 * (R/:2607-2720, call R/:4135) -> clamp and commit (R/:4150-4181).  Ten launches per step on the context's stream, no host round trip.
 * Field order: [0] = P (vegetation), [1] = G (grazer), [2] = Pr (predator), [3] = W (soil water), [4] = O (surface water).
 * The three means come from a fixed summation tree (a few ulp from the reference's left-to-right sums); given them, every output
 * equals the CPU oracle bit for bit.
 */
#ifndef TM_CONSUMER3_H
#define TM_CONSUMER3_H

#include <stdint.h>
#include "tm_consumer.h"   /* tm_graz_params */

#ifdef __cplusplus
extern "C" {
#endif

/* Dw, Do are D[3]/dx^2 and D[4]/dx^2 (Dxd2[s], R/:3463-3480) */
typedef struct tm_cons3_params {
    int32_t L, pad_;
    double dt, a, c, gmax, alpha, rW, W0, K1, K2, Dw, Do;
    double A01, H01, E1;       /* A[0][1], H[0][1], E[1]: grazer on vegetation */
    double A12, H12, E2;       /* A[1][2], H[1][2], E[2]: predator on grazer   */
} tm_cons3_params;

typedef struct tm_cons3_ctx tm_cons3_ctx;

int tm_cons3_create(int L, int device, tm_cons3_ctx **out);
int tm_cons3_destroy(tm_cons3_ctx *);
const char *tm_cons3_last_error(const tm_cons3_ctx *);   /* ctx == NULL: error of the last failed create on this thread */

/* veg: the vegetation's constants as the SPB = 1 step takes them (struct tm_params of tm_lattice.h: L, dt, dx, d, D0, sigma0, clamp_eps,
 * clamp_lo are read); graz, pred: D[s], v[s], M[s], sigma[s] of the two consumers and D0_floor = D[0]/dx^2 (R/:4014); central: the full
 * distance-sorted disc list computeNeighboringSitesCentral(int(r_max_effective)) as [n_central][2]; fr_graz, fr_pred = pR[s] /
 * (dx r_max_effective) (R/:3409). */
struct tm_params;
int tm_cons3_setup(tm_cons3_ctx *, const struct tm_params *veg, const tm_graz_params *graz, const tm_graz_params *pred,
                   const tm_cons3_params *pc, const int32_t *central, int n_central, double fr_graz, double fr_pred);
/* Rietkerk_FPE flavour of the step: the time-loop body of rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=3,
 * Rietkerk_FPE/rietkerk_bjork_FKE.cpp:4094-4462 -- calc_gamma_vel_3Sp_NonRefugia (:3371-3483: per site one uniform per species for a new
 * advection direction, in the order of decreasing perception radius, the vegetation's draw included; the perception fields ACCUMULATED
 * onto their old values; a vanished consumer gets gamma = 1) -> vegetation Dornic -> five-field RK4 over f_DorFKE_3Sp (:2611-2650: both
 * mortalities in the reaction terms; no consumer has a Dornic update) -> advdiff_FKE_MultiSp on grazer and predator (:4308, the kernels of
 * tm_fke.h on the resident fields) -> clamp with the veg->clamp_eps / clamp_lo given to setup (1e-10 and 0 in the reference, :4458).
 * t[2]: movement windows and stencil sizes of grazer and predator (struct tm_fke_terms, CPU-path convention); gauss_D[2], gauss_V[2]:
 * their stencils (host).  The directions' Philox streams use the seed xor 0xA0761D6478BD642F; in tm_cons3_replay the second array
 * carries them (cell i, k-th species in radius order: u[i*stride + k]) and the third is ignored.  Eighteen launches per step.
 * t == NULL switches back to the Rietkerk_FD flavour. */
struct tm_fke_terms;
int tm_cons3_set_fpe(tm_cons3_ctx *, const struct tm_fke_terms *t, const double *const *gauss_D, const double *const *gauss_V, double M1, double M2);
/* fields [5][L*L]; gamma [2][L*L] (grazer, predator) or NULL (zeros; the three-species perception function assigns, so the entry
 * values only matter for a species it skips) */
/* 1: the three spatial means by left-to-right sums as mean_of_vector computes them (R/:1211-1219; see tm_cons_step2_set_strict_means) */
int tm_cons3_set_strict_means(tm_cons3_ctx *, int on);
int tm_cons3_upload(tm_cons3_ctx *, const double *fields, const double *gamma);
/* nsteps steps with Philox streams keyed (seed, trial, step0 + s, cell); the grazer's and the predator's streams use the seed xor
 * 0x9E3779B97F4A7C15 and 0xD1B54A32D192ED03.  Asynchronous. */
int tm_cons3_run(tm_cons3_ctx *, uint64_t seed, uint32_t trial, uint32_t step0, int nsteps);
/* ONE step with host-supplied uniforms: cell i of species s takes its k-th uniform from u_s[i*stride + k] */
int tm_cons3_replay(tm_cons3_ctx *, const double *u_veg, const double *u_graz, const double *u_pred, uint32_t stride);
/* any of: fields [5][L*L], gamma [2][L*L], v_eff [2][L*L][2] (the last step's advection vectors), means[6] = {<P>, <G>, <Pr>, #{P > 0},
 * disc offsets summed by grazer, by predator} of the state the LAST step started from, overflow (replay draws past `stride`) */
int tm_cons3_download(tm_cons3_ctx *, double *fields, double *gamma, double *v_eff, double *means, uint64_t *overflow);
int tm_cons3_last_ms(tm_cons3_ctx *, float *ms, int *n_launches);

#ifdef __cplusplus
}
#endif
#endif
