/*
 * oracle/ref_harness_fke.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference FKE movement code
 * (/root/reference/simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp, SPB=2, CPU path, compiled where it
 * lies by oracle/Makefile into oracle/_ref/): set_global_FKE_params (:195), precompute_gaussian (:1735),
 * advdiff_FKE_MultiSp (:3614), and the replicate runner rietkerk_DorFPE_2D_MultiSp (:3697) under sequential replay.
 * Everything numerical is the reference's own code; this file only moves flat arrays in and out of its std::vector
 * containers and reads back its inline globals.  Three symbols of the reference object (save_frame, save_prelimframe and the
 * frame initialiser the runner calls as shipped, init_gaussframe :3962) are weakened by oracle/Makefile and replaced below, as
 * ref_harness_riet2.cpp does for Rietkerk_FD.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FPE */
#include <fcntl.h>

std::vector<double> precompute_gaussian(double sigma, int Rbound); /* :1735, not declared in the header */

namespace {
struct quiet {
    int o, e, n;
    quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
              n = open("/dev/null", O_WRONLY); o = dup(1); e = dup(2); dup2(n, 1); dup2(n, 2); }
    ~quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
               dup2(o, 1); dup2(e, 2); close(o); close(e); close(n); }
};
thread_local const double *g_init = nullptr;   /* Sp*L*L initial frame */
thread_local double *g_frames = nullptr;       /* captured frames [cap][Sp*L*L] */
thread_local double *g_frame_t = nullptr;
thread_local int g_frames_cap = 0, g_frames_n = 0;
}

/* --- link-time replacements ------------------------------------------------------------------------------------------- */
void init_gaussframe(D2Vec_Double &array, int size, vector<double> &, vector<double> &, int)                  /* replaces :1089 */
{
    for (int s = 0; s < Sp; s++)
        for (int i = 0; i < size; i++) array[s][i] = g_init[(size_t)s * size + i];
}
void save_frame(D2Vec_Double &Rho_t, const string &, const string &, double, double, double,
                double t, double, double, double, int, int g, string, bool)                                  /* replaces :2829 */
{
    if ((int)Rho_t.size() != Sp || !g_frames || g_frames_n >= g_frames_cap) return;   /* gamma frames (SpB rows) are not captured */
    const size_t n = (size_t)g * g;
    for (int s = 0; s < Sp; s++)
        for (size_t i = 0; i < n; i++) g_frames[((size_t)g_frames_n * Sp + s) * n + i] = Rho_t[s][i];
    g_frame_t[g_frames_n++] = t;
}
void save_prelimframe(D2Vec_Double &, const string &, const string &, double, double, double,
                      double, double, double, double, int, int, string, bool, bool) {}                      /* replaces :2723 */

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

/* One replicate-runner call: rietkerk_DorFPE_2D_MultiSp (:3697) compiled with SPB = 2 (CPU movement path), one replicate, under
 * sequential replay.  par = {a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, sigma0, sigma1, v1, aij, hij, ej, m1, pR1, dt, dx,
 * dtAdv1} (the runner vector of make_runner2_golden.py plus the grazer's advection time scale); set_global_FKE_params is called as
 * the driver does (order_2stoc_unity_rietkerk.cpp:202, CI = 4, Dt_analytic = dt).  Frames (all four fields, full precision) are
 * captured at t == 0 and at every frame_t window.  Returns the number of frames captured. */
int ref_fke_run(const double *par, int L, const double *init, const double *t_meas_in, int n_tmeas, const double *frame_t_in,
                int n_frame_t, const double *u, uint64_t n_u, double *frames_out, double *frame_times_out, int frames_cap,
                uint64_t *consumed, uint64_t *overflow)
{
    double D[Sp] = {par[9], par[10], par[11], par[12]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[Sp] = {par[13], par[14], 0, 0};
    double v[SpB] = {0, par[15]};
    double A[SpB][SpB] = {{0, par[16]}, {par[16], 0}}, H[SpB][SpB] = {{0, par[17]}, {par[17], 0}};
    double E[SpB] = {1.0, par[18]}, M[SpB] = {par[6], par[19]}, pR[Sp] = {0.0, par[20]};
    double dtAdv[SpB] = {0.0, par[23]};
    int dtV[SpB] = {0, 1};
    double clow[2 * Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
    const double dt = par[21], dx = par[22];
    set_global_FKE_params(D, v, dtAdv, dt, dx, L, 4);
    frame_tmeas.assign(frame_t_in, frame_t_in + n_frame_t);
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    D2Vec_Double Rho;
    g_init = init; g_frames = frames_out; g_frame_t = frame_times_out; g_frames_cap = frames_cap; g_frames_n = 0;
    tm_replay_arm(u, n_u);
    {
        quiet q;
        rietkerk_DorFPE_2D_MultiSp(Rho, t_meas, t_meas.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma,
                                   par[0], par[0], 1.0 / 24.0, A, H, E, M, pR, dtV, clow, dt, dx, 10000.0, 1, L);
    }
    if (consumed) *consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    g_init = nullptr; g_frames = nullptr;
    return g_frames_n;
}

/* sorted disc offsets of radius R (computeNeighboringSitesCentral :1686): out[cap][2]; returns their number (-1: cap too small) */
int ref_fke_central_sites(int R, int *out, int cap)
{
    const std::vector<std::pair<int, int>> v = computeNeighboringSitesCentral(R);
    if ((int)v.size() > cap) return -1;
    for (size_t k = 0; k < v.size(); k++) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
    return (int)v.size();
}
}
