/*
 * oracle/ref_harness_riet2.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference Rietkerk integrator compiled with -DSPB=2 (vegetation + grazer;
 * /root/reference/simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp, compiled where it lies by oracle/Makefile into
 * oracle/_ref/): computeNeighboringSitesCentral (:1661), calc_gamma_2Sp_NonRefugia (:2530),
 * RK4_Integrate_Stochastic_2Sp (:2472) with f_2Dor_2Sp (:2445), determine_neighbours_Sq4 (:1609).
 * Everything numerical is the reference's own code; this file only moves flat arrays in and out of its containers.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FD */
#include <fcntl.h>

/* definitions in rietkerk_bjork_basic.cpp that the header does not declare (or declares with stale signatures) */
void calc_gamma_2Sp_NonRefugia(const vector<pair<int, int>> &centralNeighboringSites, D2Vec_Double &Rho_t, D2Vec_Double &gamma,
                               double (&Rho_avg)[Sp], vector<std::pair<double, int>> &rfrac, double nVeg_frac, int r_max, int L); /* :2530 */
void RK4_Integrate_Stochastic_2Sp(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &K1, D2Vec_Double &K2, D2Vec_Double &K3,
                                  D2Vec_Double &K4, D3Vec_Int &nR2, double a, double c, double gmax, double alpha, double rW, double W0,
                                  double (&Dxd2)[Sp], double (&K)[3], double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB],
                                  double t, double dt, double dx, int g);                                                           /* :2472 */
void set_global_system_params(double dt, double dx);                                                                               /* :165 */

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

/* --- link-time replacements (the reference object's symbols are weakened by oracle/Makefile, as for SPB = 1) -------- */
void init_exprtk_randbiMFTframe(D2Vec_Double &array, int size, double, double, double, double, double[])   /* replaces :396 */
{
    for (int s = 0; s < Sp; s++)
        for (int i = 0; i < size; i++) array[s][i] = g_init[(size_t)s * size + i];
}
void save_frame(D2Vec_Double &Rho_t, const string &, const string &, double, double, double,
                double t, double, double, double, int, int g, string, bool)                                  /* replaces :2836 */
{
    if ((int)Rho_t.size() != Sp || !g_frames || g_frames_n >= g_frames_cap) return;   /* gamma frames (SpB rows) are not captured */
    const size_t n = (size_t)g * g;
    for (int s = 0; s < Sp; s++)
        for (size_t i = 0; i < n; i++) g_frames[((size_t)g_frames_n * Sp + s) * n + i] = Rho_t[s][i];
    g_frame_t[g_frames_n++] = t;
}
void save_prelimframe(D2Vec_Double &, const string &, const string &, double, double, double,
                      double, double, double, double, int, int, string, bool, bool) {}                      /* replaces :2720 */

extern "C" {

int ref2_spb(void) { return SpB; }

/* One replicate runner call: rietkerk_Dornic_2D_MultiSp (:3397) compiled with SPB = 2, one replicate, under sequential replay.
 * par = {a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, sigma0, sigma1, v1, aij, hij, ej, m1, pR1, dt, dx}.
 * Frames (all four fields, full precision) are captured at t == 0 and at every frame_t window.  Returns frames captured. */
int ref2_run(const double *par, int L, const double *init, const double *t_meas_in, int n_tmeas, const double *frame_t_in,
             int n_frame_t, const double *u, uint64_t n_u, double *frames_out, double *frame_times_out, int frames_cap,
             uint64_t *consumed, uint64_t *overflow)
{
    double D[Sp] = {par[9], par[10], par[11], par[12]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[Sp] = {par[13], par[14], 0, 0};
    double v[SpB] = {0, par[15]};
    double A[SpB][SpB] = {{0, par[16]}, {par[16], 0}}, H[SpB][SpB] = {{0, par[17]}, {par[17], 0}};
    double E[SpB] = {1.0, par[18]}, M[SpB] = {par[6], par[19]}, pR[Sp] = {0.0, par[20]};
    int dtV[SpB] = {0, 1};
    double clow[2 * Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
    const double dt = par[21], dx = par[22];
    set_global_system_params(dt, dx);
    frame_tmeas.assign(frame_t_in, frame_t_in + n_frame_t);
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    D2Vec_Double Rho;
    g_init = init; g_frames = frames_out; g_frame_t = frame_times_out; g_frames_cap = frames_cap; g_frames_n = 0;
    tm_replay_arm(u, n_u);
    {
        quiet q;
        rietkerk_Dornic_2D_MultiSp(Rho, t_meas, t_meas.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma,
                                   par[0], par[0], 1.0 / 24.0, A, H, E, M, pR, dtV, clow, dt, dx, 10000.0, 1, L);
    }
    if (consumed) *consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    g_init = nullptr; g_frames = nullptr;
    return g_frames_n;
}

/* sorted disc offsets of radius R: out[cap][2]; returns their number (or -1 when cap is too small) */
int ref2_central_sites(int R, int *out, int cap)
{
    const std::vector<std::pair<int, int>> v = computeNeighboringSitesCentral(R);
    if ((int)v.size() > cap) return -1;
    for (size_t k = 0; k < v.size(); k++) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
    return (int)v.size();
}

/* gamma of the grazer: veg, graz [L*L]; gamma in/out [L*L]; rho_avg[2] = {<veg>, <grazer>}; fr = grazer's radius fraction */
void ref2_calc_gamma(int L, int R, const double *veg, const double *graz, double *gamma, const double *rho_avg, double fr, double nVeg_frac)
{
    const size_t n = (size_t)L * L;
    const std::vector<std::pair<int, int>> central = computeNeighboringSitesCentral(R);
    D2Vec_Double Rho_t(Sp, std::vector<double>(n, 0.0)), gam(SpB, std::vector<double>(n, 0.0));
    for (size_t i = 0; i < n; i++) { Rho_t[0][i] = veg[i]; Rho_t[1][i] = graz[i]; gam[1][i] = gamma[i]; }
    double Rho_avg[Sp] = {rho_avg[0], rho_avg[1], 0.0, 0.0};
    std::vector<std::pair<double, int>> rfrac = {{fr, 1}, {0.0, 0}};   /* sorted in descending order, R/:3411-3415 */
    calc_gamma_2Sp_NonRefugia(central, Rho_t, gam, Rho_avg, rfrac, nVeg_frac, R, L);
    for (size_t i = 0; i < n; i++) gamma[i] = gam[1][i];
}

/* one RK4 step on fields [4][L*L] = P, G, W, O; par = {dt, dx, a, c, gmax, alpha, rW, W0, K1, K2, D2, D3, A01, H01, E1} */
void ref2_rk4_2sp(int L, const double *par, double *fields)
{
    const size_t n = (size_t)L * L;
    const double dt = par[0], dx = par[1];
    set_global_system_params(dt, dx);
    D2Vec_Double Rho_t(Sp, std::vector<double>(n)), Rho_tsar(Sp, std::vector<double>(n, 0.0));
    D2Vec_Double K1(Sp, std::vector<double>(n, 0.0)), K2(K1), K3(K1), K4(K1);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) Rho_t[s][i] = fields[s * n + i];
    D3Vec_Int nR2(n, std::vector<std::vector<int>>(2, std::vector<int>(2, 0)));
    determine_neighbours_Sq4(L, nR2);
    double Dxd2[Sp] = {0.0, 0.0, par[10] / (dx * dx), par[11] / (dx * dx)};       /* R/:3463-3480: D[s]/dx2 */
    double K[3] = {0.0, par[8], par[9]};
    double A[SpB][SpB] = {{0, par[12]}, {par[12], 0}}, H[SpB][SpB] = {{0, par[13]}, {par[13], 0}}, E[SpB] = {1.0, par[14]};
    RK4_Integrate_Stochastic_2Sp(Rho_t, Rho_tsar, K1, K2, K3, K4, nR2, par[2], par[3], par[4], par[5], par[6], par[7], Dxd2, K, A, H, E,
                                 0.0, dt, dx, L);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) fields[s * n + i] = Rho_t[s][i];
}
}
