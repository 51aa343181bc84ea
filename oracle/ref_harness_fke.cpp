/*
 * oracle/ref_harness_fke.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference FKE movement code
 * (/root/reference/simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp, SPB=2, CPU path, compiled where it
 * lies by oracle/Makefile into oracle/_ref/): set_global_FKE_params (:195), precompute_gaussian (:1735),
 * advdiff_FKE_MultiSp (:3614).  Everything numerical is the reference's own code; this file only moves
 * flat arrays in and out of its std::vector containers and reads back its inline globals.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FPE */

std::vector<double> precompute_gaussian(double sigma, int Rbound); /* :1735, not declared in the header */

extern "C" {

int ref_fke_spb(void) { return SpB; }

/* D[Sp], v[SpB], dtAdv[SpB] as the drivers pass them (order_2stoc_unity_rietkerk.cpp:202).
 * out[8] = {beta_vel, sigma_D, sigma_vD.first, mu_vel_prefactor, bound_D, bound_vD.first, bound_vD.second, epsilon} of species 1. */
void ref_fke_set_params(const double *D_, const double *v_, const double *dtAdv_, double Dt, double dx, int L, int CI, double *out)
{
    double D[Sp], v[SpB], dtAdv[SpB];
    for (int s = 0; s < Sp; s++) D[s] = D_[s];
    for (int s = 0; s < SpB; s++) { v[s] = v_[s]; dtAdv[s] = dtAdv_[s]; }
    set_global_FKE_params(D, v, dtAdv, Dt, dx, L, CI);
    out[0] = beta_vel[1]; out[1] = sigma_D[1]; out[2] = sigma_vD[1].first; out[3] = mu_vel_prefactor[1];
    out[4] = sig_D_ScaledBounds[1]; out[5] = sig_vD_ScaledBounds[1].first; out[6] = sig_vD_ScaledBounds[1].second;
    out[7] = epsilon;
}

/* overrides of the inline globals advdiff_FKE_MultiSp reads (to exercise wider windows than the drivers' constants give) */
void ref_fke_override(int bound_D, int v_first, int v_second, double prefactor, int L)
{
    sig_D_ScaledBounds[1] = bound_D;
    sig_vD_ScaledBounds[1] = {v_first, v_second};
    mu_vel_prefactor[1] = prefactor;
    L2 = (double)L * L;
}

void ref_fke_precompute_gaussian(double sigma, int Rbound, double *out)
{
    const std::vector<double> s = precompute_gaussian(sigma, Rbound);
    for (int x = 0; x <= Rbound; x++) out[x] = s[x];
}

/* one movement step of species 1 (the grazer): rho, gamma [g*g]; v_eff [g*g][2]; gD [nD]; gV [nV]; out [g*g] */
void ref_fke_advdiff(int g, const double *rho, const double *gamma, const double *v_eff, const double *gD, int nD,
                     const double *gV, int nV, double *out)
{
    const size_t n = (size_t)g * g;
    D2Vec_Double Rho_t(SpB, std::vector<double>(n, 0.0)), Rho_tsar(SpB, std::vector<double>(n, 0.0)), gam(SpB, std::vector<double>(n, 0.0));
    D2Vec_Pair_Double veff(SpB, std::vector<std::pair<double, double>>(n, {0.0, 0.0}));
    D2Vec_Double sD(SpB, std::vector<double>(1, 0.0)), sV(SpB, std::vector<double>(1, 0.0));
    for (size_t i = 0; i < n; i++) { Rho_t[1][i] = rho[i]; gam[1][i] = gamma[i]; veff[1][i] = {v_eff[2 * i], v_eff[2 * i + 1]}; }
    sD[1].assign(gD, gD + nD); sV[1].assign(gV, gV + nV);
    advdiff_FKE_MultiSp(Rho_t, Rho_tsar, gam, veff, sD, sV, g);
    for (size_t i = 0; i < n; i++) out[i] = Rho_t[1][i];   /* the reference copies Rho_tsar back into Rho_t, :3690-3691 */
}
}
