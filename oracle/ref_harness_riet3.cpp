/*
 * oracle/ref_harness_riet3.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference Rietkerk integrator compiled with -DSPB=3 (vegetation + grazer + predator;
 * /root/reference/simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp, compiled where it lies by oracle/Makefile into oracle/_ref/):
 * the replicate runner rietkerk_Dornic_2D_MultiSp (:3397) under sequential replay, calc_gamma_3Sp_NonRefugia (:3153) and
 * RK4_Integrate_Stochastic_MultiSp (:2652) with f_2Dor_3Sp (:2607).  Everything numerical is the reference's own code; this file
 * only moves flat arrays in and out of its containers.  Field order: P, G, Pr, W, O.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FD */
#include <fcntl.h>

void calc_gamma_3Sp_NonRefugia(const vector<pair<int, int>> &centralNeighboringSites, D2Vec_Double &Rho_t, D2Vec_Double &gamma,
                               double (&Rho_avg)[Sp], vector<std::pair<double, int>> &rfrac, vector<std::pair<int, int>> &dtV_counter,
                               double nVeg_frac, int r_max, int L);                                                                /* :3153 */
void RK4_Integrate_Stochastic_MultiSp(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &K1, D2Vec_Double &K2, D2Vec_Double &K3,
                                      D2Vec_Double &K4, D3Vec_Int &nR2, double a, double c, double gmax, double alpha, double rW, double W0,
                                      double (&Dxd2)[Sp], double (&K)[3], double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB],
                                      double t, double dt, double dx, int g);                                                       /* :2652 */
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

/* --- link-time replacements (the reference object's symbols are weakened by oracle/Makefile) ---------------------------------- */
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

int ref3_spb(void) { return SpB; }

/* par = {a, c, gmax, alpha, rW, W0, d, K1, K2,  D0, D1, D2, D3, D4,  sigma0, sigma1, sigma2,  v1, v2,  aij, hij, ej,  ajm, hjm, em,
 *        m1, m2,  pR1, pR2,  dt, dx}  (31 entries; D3, D4: soil and surface water) */
static void unpack(const double *par, double (&D)[Sp], double (&K)[3], double (&sigma)[Sp], double (&v)[SpB], double (&A)[SpB][SpB],
                   double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double (&pR)[Sp])
{
    for (int s = 0; s < Sp; s++) D[s] = par[9 + s];
    K[0] = 0.0; K[1] = par[7]; K[2] = par[8];
    sigma[0] = par[14]; sigma[1] = par[15]; sigma[2] = par[16]; sigma[3] = 0; sigma[4] = 0;
    v[0] = 0; v[1] = par[17]; v[2] = par[18];
    for (int i = 0; i < SpB; i++) for (int j = 0; j < SpB; j++) { A[i][j] = 0; H[i][j] = 0; }
    A[0][1] = A[1][0] = par[19]; H[0][1] = H[1][0] = par[20]; A[1][2] = A[2][1] = par[22]; H[1][2] = H[2][1] = par[23];
    E[0] = 1.0; E[1] = par[21]; E[2] = par[24];
    M[0] = par[6]; M[1] = par[25]; M[2] = par[26];
    for (int s = 0; s < Sp; s++) pR[s] = 0.0;
    pR[1] = par[27]; pR[2] = par[28];
}

/* One replicate-runner call under sequential replay; frames (all five fields, full precision) at t == 0 and at every frame_t window. */
int ref3_run(const double *par, int L, const double *init, const double *t_meas_in, int n_tmeas, const double *frame_t_in,
             int n_frame_t, const double *u, uint64_t n_u, double *frames_out, double *frame_times_out, int frames_cap,
             uint64_t *consumed, uint64_t *overflow)
{
    double D[Sp], K[3], sigma[Sp], v[SpB], A[SpB][SpB], H[SpB][SpB], E[SpB], M[SpB], pR[Sp];
    unpack(par, D, K, sigma, v, A, H, E, M, pR);
    int dtV[SpB] = {0, 1, 1};
    double clow[2 * Sp] = {0, 1, 1, 1, 1, 5, 1, 1, 1, 1};
    const double dt = par[29], dx = par[30];
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

int ref3_central_sites(int R, int *out, int cap)
{
    const std::vector<std::pair<int, int>> v = computeNeighboringSitesCentral(R);
    if ((int)v.size() > cap) return -1;
    for (size_t k = 0; k < v.size(); k++) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
    return (int)v.size();
}

/* gamma of grazer and predator: fields [3][L*L] = P, G, Pr; gamma [2][L*L] out; rho_avg[3]; fr[2] = the two radius fractions */
void ref3_calc_gamma(int L, int R, const double *fields, double *gamma, const double *rho_avg, const double *fr, double nVeg_frac)
{
    const size_t n = (size_t)L * L;
    const std::vector<std::pair<int, int>> central = computeNeighboringSitesCentral(R);
    D2Vec_Double Rho_t(Sp, std::vector<double>(n, 0.0)), gam(SpB, std::vector<double>(n, 0.0));
    for (int s = 0; s < 3; s++) for (size_t i = 0; i < n; i++) Rho_t[s][i] = fields[s * n + i];
    for (int s = 1; s < 3; s++) for (size_t i = 0; i < n; i++) gam[s][i] = gamma[(s - 1) * n + i];
    double Rho_avg[Sp] = {rho_avg[0], rho_avg[1], rho_avg[2], 0.0, 0.0};
    std::vector<std::pair<double, int>> rfrac = {{fr[0], 1}, {fr[1], 2}, {0.0, 0}};
    sort(rfrac.begin(), rfrac.end(), [](std::pair<double, int> &a, std::pair<double, int> &b) { return a.first > b.first; });   /* :3411-3415 */
    std::vector<std::pair<int, int>> dtV_counter(SpB, {1, 0});
    calc_gamma_3Sp_NonRefugia(central, Rho_t, gam, Rho_avg, rfrac, dtV_counter, nVeg_frac, R, L);
    for (int s = 1; s < 3; s++) for (size_t i = 0; i < n; i++) gamma[(s - 1) * n + i] = gam[s][i];
}

/* one RK4 step on fields [5][L*L]; par = {dt, dx, a, c, gmax, alpha, rW, W0, K1, K2, D3, D4, A01, H01, E1, A12, H12, E2} */
void ref3_rk4(int L, const double *par, double *fields)
{
    const size_t n = (size_t)L * L;
    const double dt = par[0], dx = par[1];
    set_global_system_params(dt, dx);
    D2Vec_Double Rho_t(Sp, std::vector<double>(n)), Rho_tsar(Sp, std::vector<double>(n, 0.0));
    D2Vec_Double K1(Sp, std::vector<double>(n, 0.0)), K2(K1), K3(K1), K4(K1);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) Rho_t[s][i] = fields[s * n + i];
    D3Vec_Int nR2(n, std::vector<std::vector<int>>(2, std::vector<int>(2, 0)));
    determine_neighbours_Sq4(L, nR2);
    double Dxd2[Sp] = {0.0, 0.0, 0.0, par[10] / (dx * dx), par[11] / (dx * dx)};
    double K[3] = {0.0, par[8], par[9]};
    double A[SpB][SpB] = {{0, par[12], 0}, {par[12], 0, par[15]}, {0, par[15], 0}};
    double H[SpB][SpB] = {{0, par[13], 0}, {par[13], 0, par[16]}, {0, par[16], 0}};
    double E[SpB] = {1.0, par[14], par[17]};
    RK4_Integrate_Stochastic_MultiSp(Rho_t, Rho_tsar, K1, K2, K3, K4, nR2, par[2], par[3], par[4], par[5], par[6], par[7], Dxd2, K, A, H, E,
                                     0.0, dt, dx, L);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) fields[s * n + i] = Rho_t[s][i];
}
}
