/*
 * oracle/tm_oracle_fke.c -- CPU ORACLE of the FKE movement step.  TEST INFRASTRUCTURE, NOT PRODUCT
 * (see tm_oracle.h for who may load this library).
 *
 * Restates, in the reference's own loop order (a scatter over source sites, accumulating into
 * Rho_tsar left to right), the analytic Fokker-Planck movement of one consumer species:
 *   advdiff_FKE_MultiSp      F/:3614-3695  (CPU path)
 *   updateMovement           simulations/Rietkerk_FPE/kernel_advdiff_FKE2D.cu:92-180 (window convention only)
 *   precompute_gaussian      F/:1735-1749
 *   set_global_FKE_params    F/:195-262    (per species s >= 1)
 * F/ = simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp.
 *
 * Parity status: PINNED -- tests/golden/fke_move.npz holds outputs of the reference's own
 * advdiff_FKE_MultiSp / precompute_gaussian / set_global_FKE_params compiled from
 * /root/reference (oracle/Makefile `ref`, oracle/ref_harness_fke.cpp, SPB=2).
 */
#include "tm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* the reference's PI: 15 terms of the Bailey-Borwein-Plouffe series summed upwards in double arithmetic
 * (CalculatePi<14>, Rietkerk_FPE/rietkerk_bjork_basic.h:241-259, :284) */
static double tmo_pi(void)
{
    double pi = 0.0, p16 = 1.0;
    for (int n = 0; n <= 14; n++) {
        pi = (1.0 / p16 * (4.0 / (8 * n + 1.0) - 2.0 / (8 * n + 4.0) - 1.0 / (8 * n + 5.0) - 1.0 / (8 * n + 6.0))) + pi;
        p16 *= 16.0;
    }
    return pi;
}

/* F/:1735-1749.  out[Rbound+1]. */
void tmo_fke_precompute_gaussian(double sigma, int Rbound, double *out)
{
    double sum = 0.0;
    const double factor = 1.0 / (sqrt(2 * tmo_pi()) * sigma);
    for (int x = 0; x <= Rbound; ++x) {
        out[x] = factor * exp(-0.5 * (x * x) / (sigma * sigma));
        sum += (x > 0) ? 2 * out[x] : out[0];
    }
    for (int x = 0; x <= Rbound; ++x) out[x] /= sum;
}

/* F/:195-262 for one species: D = D[s], v = v[s], dtAdv = dtAdv[s] (only read under MV_INVARIANCE, which
 * the drivers do not define), Dt = Dt_Analytic (already the smallest positive dtAdv). */
void tmo_fke_params(double D, double v, double Dt, double dx, int CI, tmo_fke_derived *o)
{
    const double beta_vel = 3.0 / Dt;                                             /* F/:229 */
    const double exp_beta_Dt = exp(-beta_vel * Dt), exp_2beta_Dt = exp(-2 * beta_vel * Dt); /* F/:231 */
    const double sigma_v = sqrt(2 * beta_vel * 1 * fabs(v));                      /* F/:233 */
    const double sigma_D = sqrt(2 * D * Dt);                                      /* F/:234 */
    const double sigma_vx = (sigma_v / beta_vel) * sqrt((Dt + (1 / (2 * beta_vel)) * (1 - exp_2beta_Dt)
                                                         - (2 / beta_vel) * (1 - exp_beta_Dt))); /* F/:235-236 */
    o->beta_vel = beta_vel;
    o->sigma_D = sigma_D;
    o->sigma_vD = sigma_vx;
    o->bound_D = (int)(CI * sigma_D / dx);                                        /* F/:244 */
    o->bound_vD_first = (int)(CI * sigma_vx / dx);                                /* F/:246 (isotropic: .second is the same) */
    o->bound_vD_second = o->bound_vD_first;
    o->mu_vel_prefactor = (1 / dx) * (Dt - ((1 - exp_beta_Dt) / beta_vel));       /* F/:249 */
}

/* F/:3614-3695 for one species.  rho, gamma: [g*g]; v_eff: [g*g][2] = {first, second}; gD: [bound_D+1];
 * gV: [max window half-width + 1]; out: [g*g], zeroed here (the caller does it in the reference, F/:4303-4307).
 * Windows: rows [mu_x - v_row_lo, mu_x + v_row_hi], columns [mu_y - v_col_lo, mu_y + v_col_hi]:
 *   CPU path F/:3640-3641      lo = sig_vD_ScaledBounds[s].first, hi = .second on BOTH axes
 *   CUDA kernel .cu:126-127    rows +-.x, columns +-.y */
int tmo_fke_move(const tmo_fke_terms *t, const double *rho, const double *gamma, const double *v_eff,
                 const double *gD, const double *gV, double *out)
{
    const int g = t->L;
    if (g < 1 || t->bound_D < 0) return -1;
    const long L2 = (long)g * g;
    memset(out, 0, sizeof(double) * (size_t)L2);
    for (long i = 0; i < L2; i++) {
        if (rho[i] < t->epsilon) continue;                                        /* F/:3622 */
        const int c_i = (int)(i / g), c_j = (int)(i % g);
        const double gamma2 = gamma[i] * gamma[i];
        const int min_p = c_i - t->bound_D, max_p = c_i + t->bound_D;
        const int min_q = c_j - t->bound_D, max_q = c_j + t->bound_D;
        const int mu_x = (int)(round(c_i + v_eff[2 * i] * t->mu_vel_prefactor) + g) % g;       /* F/:3635 */
        const int mu_y = (int)(round(c_j + v_eff[2 * i + 1] * t->mu_vel_prefactor) + g) % g;   /* F/:3637 */
        const int min_p_vx = mu_x - t->v_row_lo, max_p_vx = mu_x + t->v_row_hi;
        const int min_q_vy = mu_y - t->v_col_lo, max_q_vy = mu_y + t->v_col_hi;
        for (int p = min_p; p <= max_p; ++p) {                                    /* F/:3645-3665 */
            const int dx = abs(p - c_i);
            for (int q = min_q; q <= max_q; ++q) {
                const int dy = abs(q - c_j);
                const long i_new = (long)((p + g) % g) * g + (q + g) % g;
                out[i_new] += gamma2 * rho[i] * (gD[dx] * gD[dy]);
            }
            for (int q = min_q_vy; q <= max_q_vy; ++q) {
                const int dy = abs(q - mu_y);
                const long i_new = (long)((p + g) % g) * g + (q + g) % g;
                out[i_new] += (gamma[i] - gamma2) * rho[i] * (gD[dx] * gV[dy]);
            }
        }
        for (int p = min_p_vx; p <= max_p_vx; ++p) {                              /* F/:3667-3687 */
            const int dx = abs(p - mu_x);
            for (int q = min_q; q <= max_q; ++q) {
                const int dy = abs(q - c_j);
                const long i_new = (long)((p + g) % g) * g + (q + g) % g;
                out[i_new] += (gamma[i] - gamma2) * rho[i] * (gV[dx] * gD[dy]);
            }
            for (int q = min_q_vy; q <= max_q_vy; ++q) {
                const int dy = abs(q - mu_y);
                const long i_new = (long)((p + g) % g) * g + (q + g) % g;
                out[i_new] += (1 - gamma[i]) * (1 - gamma[i]) * rho[i] * (gV[dx] * gV[dy]);
            }
        }
    }
    return 0;
}
