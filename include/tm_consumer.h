/*
 * include/tm_consumer.h -- C ABI of two consumer-side operators of the SPB = 2 (vegetation + grazer) Rietkerk model on B200
 * (sm_100a): the first part of SURVEY 8 f3.  Plain pointers and sizes only; no CPU fallback (tm_cons_create fails with
 * TM_E_CUDA without a device).  R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.
 *
 *   tm_cons_gamma_host   replaces calc_gamma_2Sp_NonRefugia(centralNeighboringSites, Rho_t, gamma, Rho_avg, rfrac,
 *                        nVeg_frac, r_max, L)   R/:2530-2603 (rietkerk_bjork_basic.h declares the 3-species twin at :371)
 *   tm_cons_rk4_host     replaces RK4_Integrate_Stochastic_2Sp(Rho_t, Rho_tsar, K1..K4, nR2, a, c, gmax, alpha, rW, W0,
 *                        Dxd2, K, A, H, E, t, dt, dx, g)   R/:2472-2527 with f_2Dor_2Sp R/:2445-2470
 * Not yet behind this ABI (remaining): the grazer's Dornic sampling with upwind advection (R/:3908-4114),
 * the predator (SPB = 3) variants, and the SPB >= 2 replicate runner.
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

/* Device time (ms) of the kernels of the last call and how many launches that was. */
int tm_cons_last_ms(tm_cons_ctx *, float *ms, int *n_launches);

#ifdef __cplusplus
}
#endif
#endif
