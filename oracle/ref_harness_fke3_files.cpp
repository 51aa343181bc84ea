/*
 * oracle/ref_harness_fke3_files.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * The UNMODIFIED reference runner rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE/rietkerk_bjork_FKE.cpp:3697) compiled with -DSPB=3 (CPU movement
 * path) WITH ITS OWN CSV WRITERS; only the frame initialiser the runner calls as shipped (init_gaussframe :3962) is replaced, as
 * oracle/ref_harness_fke_files.cpp does for SPB = 2.  It writes its whole output tree under a directory of the caller's choice.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FPE */
#include <fcntl.h>

namespace {
struct quiet {
    int o, e, n;
    quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
              n = open("/dev/null", O_WRONLY); o = dup(1); e = dup(2); dup2(n, 1); dup2(n, 2); }
    ~quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
               dup2(o, 1); dup2(e, 2); close(o); close(e); close(n); }
};
thread_local const double *g_init = nullptr;   /* [n_init][Sp*L*L] */
thread_local int g_n_init = 1, g_calls = 0;
}

void init_gaussframe(D2Vec_Double &array, int size, vector<double> &, vector<double> &, int)                  /* replaces :1089 */
{
    const double *f = g_init + (size_t)(g_calls % g_n_init) * Sp * size;
    for (int s = 0; s < Sp; s++)
        for (int i = 0; i < size; i++) array[s][i] = f[(size_t)s * size + i];
    g_calls++;
}

static void unpack(const double *par, double (&D)[Sp], double (&K)[3], double (&sigma)[Sp], double (&v)[SpB], double (&A)[SpB][SpB],
                   double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double (&pR)[Sp], double (&dtAdv)[SpB])
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
    dtAdv[0] = 0.0; dtAdv[1] = par[31]; dtAdv[2] = par[32];
}

extern "C" {

/* out[2][8] = {beta_vel, sigma_D, sigma_vD.first, mu_vel_prefactor, bound_D, bound_vD.first, bound_vD.second, epsilon} of species 1, 2 */
void reffke3f_set_params(const double *par, int L, double *out)
{
    double D[Sp], K[3], sigma[Sp], v[SpB], A[SpB][SpB], H[SpB][SpB], E[SpB], M[SpB], pR[Sp], dtAdv[SpB];
    unpack(par, D, K, sigma, v, A, H, E, M, pR, dtAdv);
    { quiet q; set_global_FKE_params(D, v, dtAdv, par[29], par[30], L, 4); }
    for (int s = 1; s <= 2; s++) {
        double *o = out + 8 * (s - 1);
        o[0] = beta_vel[s]; o[1] = sigma_D[s]; o[2] = sigma_vD[s].first; o[3] = mu_vel_prefactor[s];
        o[4] = sig_D_ScaledBounds[s]; o[5] = sig_vD_ScaledBounds[s].first; o[6] = sig_vD_ScaledBounds[s].second; o[7] = epsilon;
    }
}

/* par: the 33-entry vector of tests/golden/make_fpe_runner3_golden.py; the rest as reffkef_run (oracle/ref_harness_fke_files.cpp) */
uint64_t reffke3f_run(const double *par, int L, int r, const double *init, int n_init, const double *t_meas_in, int n_tmeas,
                      const double *frame_t_in, int n_frame_t, const double *u, uint64_t n_u, const char *root, const char *user_prefix,
                      double Gstar, uint64_t *overflow)
{
    double D[Sp], K[3], sigma[Sp], v[SpB], A[SpB][SpB], H[SpB][SpB], E[SpB], M[SpB], pR[Sp], dtAdv[SpB];
    unpack(par, D, K, sigma, v, A, H, E, M, pR, dtAdv);
    int dtV[SpB] = {0, 1, 1};
    double clow[2 * Sp] = {0, 1, 1, 1, 1, 5, 1, 1, 1, 1};
    const double dt = par[29], dx = par[30];
    { quiet q; set_global_FKE_params(D, v, dtAdv, dt, dx, L, 4); }
    string pre = user_prefix;
    set_Prefix(pre, 0.0, 0.0);
    const string rt = root;
    frame_folder = rt + frame_folder.substr(string("../Data/Rietkerk/").size());
    prelim_folder = rt + prelim_folder.substr(string("../Data/Rietkerk/").size());
    stat_prefix = rt + stat_prefix.substr(string("../Data/Rietkerk/").size());
    frame_tmeas.assign(frame_t_in, frame_t_in + n_frame_t);
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    D2Vec_Double Rho;
    g_init = init; g_n_init = n_init; g_calls = 0;
    tm_replay_arm(u, n_u);
    {
        quiet q;
        rietkerk_DorFPE_2D_MultiSp(Rho, t_meas, t_meas.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma,
                                   par[0], par[0], 1.0 / 24.0, A, H, E, M, pR, dtV, clow, dt, dx, 10000.0, r, L, Gstar);
    }
    const uint64_t consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    g_init = nullptr;
    return consumed;
}
}
