/*
 * oracle/tm_oracle_consumer.c -- CPU ORACLE of two consumer-side operators of the SPB = 2 (vegetation + grazer) model
 * (first part).  TEST INFRASTRUCTURE, NOT PRODUCT (see tm_oracle.h for who may load this library).
 *
 *   calc_gamma_2Sp_NonRefugia        R/:2530-2603   the grazer's perception field gamma (disc average of the vegetation)
 *   f_2Dor_2Sp                       R/:2445-2470   reaction terms of (P, G, W, O)
 *   RK4_Integrate_Stochastic_2Sp     R/:2472-2527   classical RK4 over the four fields
 *   determine_neighbours_Sq4         R/:1609-1628   periodic von-Neumann neighbours
 * R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.
 *
 * Parity status: PINNED -- tests/golden/consumer_ops.npz holds outputs of these very functions of the reference compiled
 * with -DSPB=2 from /root/reference (oracle/Makefile `ref`, oracle/ref_harness_riet2.cpp).
 */
#include "tm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* R/:2530-2603 for the grazer (s = 1, prey s-1 = vegetation).
 * offsets: the first `new_length` entries {dx, dy} of computeNeighboringSitesCentral(r_max) (R/:1661, sorted by squared
 * distance), new_length = max(1, int(nVeg_frac^2 * size)) with nVeg_frac clamped to >= 0.35 (R/:2533-2540);
 * eff = int(fr * fr * new_length) (R/:2569).  gamma is IN/OUT: the reference adds the disc sum onto the value the array
 * already holds (R/:2581 `gamma[s][i] += ...`, never reset between steps). */
static void gamma_perception(int L, const double *prey, const int32_t *offsets, int eff, double rho_avg_prey, double rho_avg_self,
                             double *gamma, double self_gone)
{
    const double eps = 1.0e-12;                                                     /* R/:2537 */
    const double rho_inverse = rho_avg_prey >= eps ? 1 / rho_avg_prey : 1 / eps;    /* R/:2548-2549 */
    for (int i = 0; i < L * L; i++) {
        const int c_i = i / L, c_j = i % L;
        if (rho_avg_self < eps) { gamma[i] = self_gone; continue; }                  /* R/:2575 (0.0); F/:2579 (1.0) */
        else if (rho_avg_prey < eps) { gamma[i] = 0.0; continue; }                   /* R/:2577 */
        for (int k = 0; k < eff; k++) {                                              /* R/:2580-2581, neighbours R/:1700-1702 */
            const int nx = (c_i + offsets[2 * k] + L) % L, ny = (c_j + offsets[2 * k + 1] + L) % L;
            gamma[i] += prey[nx * L + ny];
        }
        gamma[i] = (gamma[i] * rho_inverse) / (eff);                                 /* R/:2583 */
        if (gamma[i] > 1) gamma[i] = 1.0;                                            /* R/:2585-2588 */
        else if (gamma[i] < 0) gamma[i] = 0.0;
    }
}

void tmo_gamma_perception(int L, const double *prey, const int32_t *offsets, int eff, double rho_avg_prey, double rho_avg_self,
                          double *gamma)
{
    gamma_perception(L, prey, offsets, eff, rho_avg_prey, rho_avg_self, gamma, 0.0);
}

/* calc_gamma_vel_2Sp_NonRefugia F/:2534-2607 (Rietkerk_FPE): the same disc average; a vanished consumer gets gamma = 1 (F/:2579),
 * and the function assigns no velocities (v_eff keeps what it held). */
void tmo_gamma_perception_fpe(int L, const double *prey, const int32_t *offsets, int eff, double rho_avg_prey, double rho_avg_self,
                              double *gamma)
{
    gamma_perception(L, prey, offsets, eff, rho_avg_prey, rho_avg_self, gamma, 1.0);
}

/* f_2Dor_2Sp R/:2445-2470; fields [0]=P, [1]=G, [2]=W, [3]=O.  fpe != 0: f_DorFKE_2Sp F/:2449-2474, which differs by the grazer's
 * mortality term -M[1] G (F/:2466; in Rietkerk_FD it sits in the grazer's Dornic beta instead) */
static void f_2dor_2sp(const tmo_cons_params *p, double *f, const double *Y, int fpe, double M1)
{
    const int g = p->L, n = g * g;
    const double cgmax = p->c * p->gmax, A01H01 = p->A01 * p->H01, EA1 = p->E1 * p->A01;   /* R/:2452 */
    const double *P = Y, *G = Y + n, *W = Y + 2 * n, *O = Y + 3 * n;
    for (int i = 0; i < n; i++) {
        const int r = i / g, q = i % g;
        const int up = ((r - 1) % g + g) % g * g + q, dn = ((r + 1) % g + g) % g * g + q;   /* R/:1622-1626 */
        const int lf = r * g + ((q - 1) % g + g) % g, rt = r * g + ((q + 1) % g + g) % g;
        f[i] = cgmax * W[i] * P[i] / (W[i] + p->K1) - (p->A01 * P[i] / (1 + A01H01 * P[i])) * G[i];                     /* R/:2460 */
        f[n + i] = (EA1 * P[i] / (1 + A01H01 * P[i])) * G[i];                                                          /* R/:2462 */
        if (fpe) f[n + i] = (EA1 * P[i] / (1 + A01H01 * P[i])) * G[i] - M1 * G[i];                                     /* F/:2466 */
        f[2 * n + i] = p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i] - p->gmax * W[i] * P[i] / (W[i] + p->K1) - p->rW * W[i]
                     + (p->Dw) * (W[up] + W[dn] + W[lf] + W[rt] - 4 * W[i]);                                           /* R/:2464-2465 */
        f[3 * n + i] = p->a - p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i]
                     + (p->Do) * (O[up] + O[dn] + O[lf] + O[rt] - 4 * O[i]);                                           /* R/:2467-2468 */
    }
}

/* RK4_Integrate_Stochastic_2Sp R/:2472-2527: one step in place on f[4][L*L] */
static int rk4_2sp(const tmo_cons_params *p, double *f, int fpe, double M1);
int tmo_rk4_2sp(const tmo_cons_params *p, double *f) { return rk4_2sp(p, f, 0, 0.0); }
/* RK4_Integrate_Stochastic_2Sp of Rietkerk_FPE, F/:2476-2531, over f_DorFKE_2Sp */
int tmo_rk4_fpe_2sp(const tmo_cons_params *p, double M1, double *f) { return rk4_2sp(p, f, 1, M1); }

static int rk4_2sp(const tmo_cons_params *p, double *f, int fpe, double M1)
{
    if (!p || p->L < 1) return -1;
    const size_t n4 = (size_t)4 * p->L * p->L;
    const double dt2 = p->dt / 2.0, dt6 = p->dt / 6.0;                               /* R/:171 */
    double *K1 = malloc(sizeof(double) * n4 * 5), *K2 = K1 + n4, *K3 = K2 + n4, *K4 = K3 + n4, *Y = K4 + n4;
    if (!K1) return -1;
    f_2dor_2sp(p, K1, f, fpe, M1);
    for (size_t i = 0; i < n4; i++) Y[i] = f[i] + (dt2) * K1[i];
    f_2dor_2sp(p, K2, Y, fpe, M1);
    for (size_t i = 0; i < n4; i++) Y[i] = f[i] + (dt2) * K2[i];
    f_2dor_2sp(p, K3, Y, fpe, M1);
    for (size_t i = 0; i < n4; i++) Y[i] = f[i] + (p->dt) * K3[i];
    f_2dor_2sp(p, K4, Y, fpe, M1);
    for (size_t i = 0; i < n4; i++) f[i] += (dt6) * (K1[i] + 2.0 * K2[i] + 2.0 * K3[i] + K4[i]);   /* R/:2506 */
    free(K1);
    return 0;
}


/* ==================================================================================================================== */
/* SPB = 3: vegetation + grazer + predator                                                                              */
/* ==================================================================================================================== */
/* calc_gamma_3Sp_NonRefugia R/:3153-3268.  The caller passes new_length = max(1, int(nVeg_frac^2 * size)) (R/:3162-3163). */
void tmo_gamma_3sp(int L, const double *veg, const double *graz, const double *pred, const int32_t *offsets, int new_length,
                   double fr1, double fr2, const double *avg, double *gamma1, double *gamma2)
{
    const double eps = 1.0e-12;                                                      /* R/:3160 */
    double rho_inverse[3];
    for (int s = 0; s < 3; s++) rho_inverse[s] = avg[s] >= eps ? 1 / avg[s] : 1 / eps;   /* R/:3171-3172 */
    const double fr[3] = {0.0, fr1, fr2};
    double *gam[3] = {NULL, gamma1, gamma2};
    const double *fld[3] = {veg, graz, pred};
    for (int i = 0; i < L * L; i++) {
        const int c_i = i / L, c_j = i % L;
        for (int s = 1; s <= 2; s++) {                                               /* the order of rfrac does not matter */
            if (fr[s] == 0.0) continue;                                              /* R/:3201 */
            int eff = (int)(fr[s] * fr[s] * new_length);                             /* R/:3199 */
            if (eff < 1 && fr[s] > 0) eff = 1;                                       /* R/:3204 */
            if (avg[s] < eps) { gam[s][i] = 0.0; continue; }                         /* R/:3206 */
            else if (avg[s - 1] < eps) { gam[s][i] = 0.0; continue; }                /* R/:3208 */
            double gamma_sum = 0.0;
            const double *prey = fld[s - 1];
            if (s == 1) {
                if (avg[2] < eps) {                                                  /* no predators left, R/:3221 */
                    for (int k = 0; k < eff; k++) {
                        const int nx = (c_i + offsets[2 * k] + L) % L, ny = (c_j + offsets[2 * k + 1] + L) % L;
                        gamma_sum += prey[nx * L + ny];
                    }
                } else {
                    for (int k = 0; k < eff; k++) {                                  /* R/:3232: vegetation discounted by the predators there */
                        const int nx = (c_i + offsets[2 * k] + L) % L, ny = (c_j + offsets[2 * k + 1] + L) % L;
                        gamma_sum += prey[nx * L + ny] * (1 - pred[nx * L + ny] * rho_inverse[2]);
                    }
                }
                gam[s][i] = (gamma_sum * rho_inverse[s - 1]) / (eff);                /* R/:3239 */
            } else {
                for (int k = 0; k < eff; k++) {
                    const int nx = (c_i + offsets[2 * k] + L) % L, ny = (c_j + offsets[2 * k + 1] + L) % L;
                    gamma_sum += prey[nx * L + ny];
                }
                gam[s][i] = (gamma_sum / (eff)) * rho_inverse[s - 1];                /* R/:3248 */
            }
            if (gam[s][i] > 1) gam[s][i] = 1.0;                                      /* R/:3251-3254 */
            else if (gam[s][i] < 0) gam[s][i] = 0.0;
        }
    }
}

/* f_2Dor_3Sp R/:2607-2650.  fpe != 0: f_DorFKE_3Sp F/:2611-2650, which carries the consumers' mortalities -M[1] G, -M[2] Pr in the
 * reaction terms (F/:2636, 2639; in Rietkerk_FD they sit in the consumers' Dornic beta) */
static void f_2dor_3sp(const tmo_cons3_params *p, double *f, const double *Y, int fpe, double M1, double M2)
{
    const int g = p->L, n = g * g;
    const double cgmax = p->c * p->gmax, K2W0 = p->K2 * p->W0, A01H01 = p->A01 * p->H01, A12H12 = p->A12 * p->H12;   /* R/:2614 */
    const double EA1 = p->E1 * p->A01, EA2 = p->E2 * p->A12;                                                         /* R/:2616 */
    const double *P = Y, *G = Y + n, *R = Y + 2 * n, *W = Y + 3 * n, *O = Y + 4 * n;
    for (int i = 0; i < n; i++) {
        const int r = i / g, q = i % g;
        const int up = ((r - 1) % g + g) % g * g + q, dn = ((r + 1) % g + g) % g * g + q;   /* R/:1622-1626 */
        const int lf = r * g + ((q - 1) % g + g) % g, rt = r * g + ((q + 1) % g + g) % g;
        f[i] = cgmax * W[i] * P[i] / (W[i] + p->K1) - (p->A01 * P[i] / (1 + A01H01 * P[i])) * G[i];                     /* R/:2633 */
        f[n + i] = (EA1 * P[i] / (1 + A01H01 * P[i])) * G[i] - p->A12 * G[i] / (1 + A12H12 * G[i]) * R[i];              /* R/:2635-2636 */
        f[2 * n + i] = (EA2 * G[i] / (1 + A12H12 * G[i])) * R[i];                                                      /* R/:2638 */
        if (fpe) {
            f[n + i] = (EA1 * P[i] / (1 + A01H01 * P[i])) * G[i] - M1 * G[i] - p->A12 * G[i] / (1 + A12H12 * G[i]) * R[i];   /* F/:2636-2637 */
            f[2 * n + i] = (EA2 * G[i] / (1 + A12H12 * G[i])) * R[i] - M2 * R[i];                                        /* F/:2639 */
        }
        f[3 * n + i] = p->alpha * (P[i] + K2W0) / (P[i] + p->K2) * O[i] - p->gmax * W[i] * P[i] / (W[i] + p->K1) - p->rW * W[i]
                     + (p->Dw) * (W[up] + W[dn] + W[lf] + W[rt] - 4 * W[i]);                                           /* R/:2640-2641 */
        f[4 * n + i] = p->a - p->alpha * (P[i] + K2W0) / (P[i] + p->K2) * O[i]
                     + (p->Do) * (O[up] + O[dn] + O[lf] + O[rt] - 4 * O[i]);                                           /* R/:2643-2644 */
    }
}

/* RK4_Integrate_Stochastic_MultiSp R/:2652-2720: one step in place on f[5][L*L] */
static int rk4_3sp(const tmo_cons3_params *p, double *f, int fpe, double M1, double M2);
int tmo_rk4_3sp(const tmo_cons3_params *p, double *f) { return rk4_3sp(p, f, 0, 0.0, 0.0); }
/* RK4_Integrate_Stochastic_MultiSp of Rietkerk_FPE, F/:2652-2720, over f_DorFKE_3Sp */
int tmo_rk4_fpe_3sp(const tmo_cons3_params *p, double M1, double M2, double *f) { return rk4_3sp(p, f, 1, M1, M2); }

static int rk4_3sp(const tmo_cons3_params *p, double *f, int fpe, double M1, double M2)
{
    if (!p || p->L < 1) return -1;
    const size_t n5 = (size_t)5 * p->L * p->L;
    const double dt2 = p->dt / 2.0, dt6 = p->dt / 6.0;
    double *K1 = malloc(sizeof(double) * n5 * 5), *K2 = K1 + n5, *K3 = K2 + n5, *K4 = K3 + n5, *Y = K4 + n5;
    if (!K1) return -1;
    f_2dor_3sp(p, K1, f, fpe, M1, M2);
    for (size_t i = 0; i < n5; i++) Y[i] = f[i] + (dt2) * K1[i];
    f_2dor_3sp(p, K2, Y, fpe, M1, M2);
    for (size_t i = 0; i < n5; i++) Y[i] = f[i] + (dt2) * K2[i];
    f_2dor_3sp(p, K3, Y, fpe, M1, M2);
    for (size_t i = 0; i < n5; i++) Y[i] = f[i] + (p->dt) * K3[i];
    f_2dor_3sp(p, K4, Y, fpe, M1, M2);
    for (size_t i = 0; i < n5; i++) f[i] += (dt6) * (K1[i] + 2.0 * K2[i] + 2.0 * K3[i] + K4[i]);   /* R/:2691 */
    free(K1);
    return 0;
}
