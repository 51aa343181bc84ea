/*
 * oracle/ref_harness_riet.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference Rietkerk integrator
 * (/root/reference/simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp, SPB=1, compiled where it
 * lies by oracle/Makefile into oracle/_ref/).  The reference object's symbols are weakened with
 * objcopy, so the three definitions below take the place of the reference's frame initialiser
 * and CSV writers at link time: the runner then starts from a caller-supplied frame and hands
 * back its fields at full double precision instead of 6-digit CSV (rietkerk_bjork_basic.cpp:2863).
 * Everything numerical -- the Dornic loop (:3752-3809), RK4_Integrate_Stochastic_1Sp (:2389),
 * clamp (:4150-4181), RK4_Integrate/f_2D (:1753/:1718) -- is the reference's own code.
 */
#include "rietkerk_bjork_basic.h" /* reference header, -I/root/reference/simulations/Rietkerk_FD */
#include <fcntl.h>

/* definitions in rietkerk_bjork_basic.cpp whose header declarations are stale or missing */
void RK4_Integrate_Stochastic_1Sp(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &K1, D2Vec_Double &K2,
    D2Vec_Double &K3, D2Vec_Double &K4, D3Vec_Int &nR2, double a, double c, double gmax, double alpha, double rW,
    double W0, double (&Dxd2)[Sp], double (&K)[3], double t, double dt, double dx, int g);          /* :2389 */
void RK4_Integrate(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &K1, D2Vec_Double &K2, D2Vec_Double &K3,
    D2Vec_Double &K4, D3Vec_Int &nR2, double a, double c, double gmax, double alpha, double d, double rW, double W0,
    double (&Dxd2)[Sp], double (&K)[3], double t, double dt, double dx, int g);                     /* :1753 */
void determine_neighbours_Sq8(int g, D3Vec_Int &neighbours_Sq8);                                   /* :1631 */

namespace {
struct quiet {
    int o, e, n;
    quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
              n = open("/dev/null", O_WRONLY); o = dup(1); e = dup(2); dup2(n, 1); dup2(n, 2); }
    ~quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
               dup2(o, 1); dup2(e, 2); close(o); close(e); close(n); }
};
thread_local const double *g_init = nullptr;   /* 3*L*L initial frame */
thread_local double *g_frames = nullptr;       /* captured frames [cap][3*L*L] */
thread_local double *g_frame_t = nullptr;
thread_local int g_frames_cap = 0, g_frames_n = 0;
}

/* --- link-time replacements (strong definitions beat the weakened reference ones) --------- */

/* replaces rietkerk_bjork_basic.cpp:396 (MFT speckle frame from exprtk strings + random mask) */
void init_exprtk_randbiMFTframe(D2Vec_Double &array, int size, double, double, double, double, double[])
{
    for (int s = 0; s < Sp; s++)
        for (int i = 0; i < size; i++) array[s][i] = g_init[(size_t)s * size + i];
}

/* replaces rietkerk_bjork_basic.cpp:272 (speckle frame from mt19937_64(random_device)): the noise-off wrapper's initial frame.
 * Replicate j of a ref_riet_det_run call starts from frame j of the caller's stack. */
thread_local int g_det_replicate = 0;
void init_randconstframe(D2Vec_Double &array, int Sp_, int size, double, double[], double[])
{
    const double *f = g_init + (size_t)g_det_replicate * Sp_ * size;
    for (int s = 0; s < Sp_; s++)
        for (int i = 0; i < size; i++) array[s][i] = f[(size_t)s * size + i];
    g_det_replicate++;
}

/* replaces rietkerk_bjork_basic.cpp:2836 (6-significant-digit CSV snapshot writer) */
void save_frame(D2Vec_Double &Rho_t, const string &, const string &, double, double, double,
                double t, double, double, double, int, int g, string, bool)
{
    if ((int)Rho_t.size() != Sp || !g_frames || g_frames_n >= g_frames_cap) return;
    const size_t n = (size_t)g * g;
    for (int s = 0; s < Sp; s++)
        for (size_t i = 0; i < n; i++) g_frames[((size_t)g_frames_n * Sp + s) * n + i] = Rho_t[s][i];
    g_frame_t[g_frames_n++] = t;
}

/* replaces rietkerk_bjork_basic.cpp:2720 (per-replicate prelim CSV writer) */
void save_prelimframe(D2Vec_Double &, const string &, const string &, double, double, double,
                      double, double, double, double, int, int, string, bool, bool) {}

extern "C" {

struct ref_riet_params {
    double a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, sigma0, dt, dx;
};

/* One replicate runner call: rietkerk_Dornic_2D_MultiSp (rietkerk_bjork_basic.cpp:3397).
 * Frames are captured at t == 0 and at every frame_t window (the runner's own schedule, :3736).
 * rows_out[k] = {t, <W>, <O>, <P>, var, n_surv, n_active} (the runner's accumulator, :3444).
 * Returns the number of rows, or -1. */
int ref_riet_run(const ref_riet_params *p, int L, int r, const double *init,
                 const double *t_meas_in, int n_tmeas, const double *frame_t_in, int n_frame_t,
                 const double *u, uint64_t n_u,
                 double *frames_out, double *frame_times_out, int frames_cap, int *n_frames,
                 double *rows_out, int rows_cap, uint64_t *consumed, uint64_t *overflow)
{
    double D[Sp] = {p->D0, p->D1, p->D2};
    double K[3] = {0.0, p->K1, p->K2};
    double sigma[Sp] = {p->sigma0, 0, 0};
    double v[SpB] = {0};
    double A[SpB][SpB] = {{0}}, H[SpB][SpB] = {{0}};
    double E[SpB] = {1.0}, M[SpB] = {p->d}, pR[SpB] = {0.0};
    int dtV[SpB] = {0};
    double clow[2 * Sp] = {0, 1, 1, 5, 1, 1};
    set_global_system_params(p->dt, p->dx);
    frame_tmeas.assign(frame_t_in, frame_t_in + n_frame_t);
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    D2Vec_Double Rho;
    g_init = init; g_frames = frames_out; g_frame_t = frame_times_out;
    g_frames_cap = frames_cap; g_frames_n = 0;
    tm_replay_arm(u, n_u);
    {
        quiet q;
        rietkerk_Dornic_2D_MultiSp(Rho, t_meas, t_meas.back(), p->a, p->c, p->gmax, p->alpha, p->rW, p->W0,
            D, v, K, sigma, p->a, p->a, 1.0 / 24.0, A, H, E, M, pR, dtV, clow, p->dt, p->dx, 10000.0, r, L);
    }
    if (consumed) *consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    if (n_frames) *n_frames = g_frames_n;
    g_init = nullptr; g_frames = nullptr;
    int n = (int)Rho.size();
    if (n > rows_cap) return -1;
    for (int k = 0; k < n; k++) {
        const size_t w = Rho[k].size(); /* {a, r, L, t, <W>, <O>, <P>, var, surv, active} */
        for (int j = 0; j < 7; j++) rows_out[7 * k + j] = (3 + j < (int)w) ? Rho[k][3 + j] : 0.0;
    }
    return n;
}

/* RK4_Integrate_Stochastic_1Sp (:2389) on a caller-supplied state; f: [3][L*L] in/out. */
void ref_riet_rk4_stoch(const ref_riet_params *p, int L, double *f)
{
    const size_t n = (size_t)L * L;
    D2Vec_Double Rho_t(Sp, vector<double>(n)), tsar(Sp, vector<double>(n, 0.0)),
        K1(Sp, vector<double>(n, 0.0)), K2(K1), K3(K1), K4(K1);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) Rho_t[s][i] = f[s * n + i];
    D3Vec_Int nR2(n, D2Vec_Int(2, vector<int>(2, 0)));
    determine_neighbours_Sq4(L, nR2);
    set_global_system_params(p->dt, p->dx);
    const double dx2l = p->dx * p->dx;
    double Dxd2[Sp] = {p->D0 / dx2l, p->D1 / dx2l, p->D2 / dx2l}; /* :3463-3480 */
    double K[3] = {0.0, p->K1, p->K2};
    { quiet q;
      RK4_Integrate_Stochastic_1Sp(Rho_t, tsar, K1, K2, K3, K4, nR2, p->a, p->c, p->gmax, p->alpha,
                                   p->rW, p->W0, Dxd2, K, 0.0, p->dt, p->dx, L); }
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) f[s * n + i] = Rho_t[s][i];
}

/* nsteps of RK4_Integrate with the full RHS f_2D (:1753, :1718), as RK4_Wrapper_2D drives it
 * (:2002): the noise-off path. */
void ref_riet_rk4_det(const ref_riet_params *p, int L, int nsteps, double *f)
{
    const size_t n = (size_t)L * L;
    D2Vec_Double Rho_t(Sp, vector<double>(n)), tsar(Sp, vector<double>(n, 0.0)),
        K1(Sp, vector<double>(n, 0.0)), K2(K1), K3(K1), K4(K1);
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) Rho_t[s][i] = f[s * n + i];
    D3Vec_Int nR2(n, D2Vec_Int(3, vector<int>(3, 0)));
    determine_neighbours_Sq8(L, nR2);
    const double dx2l = p->dx * p->dx;
    double Dxd2[Sp] = {p->D0 / dx2l, p->D1 / dx2l, p->D2 / dx2l}; /* :1869-1871 */
    double K[3] = {0.0, p->K1, p->K2};
    double t = 0;
    { quiet q;
      for (int k = 0; k < nsteps; k++) {
          for (int s = 0; s < Sp; s++) tsar[s] = Rho_t[s];
          RK4_Integrate(Rho_t, tsar, K1, K2, K3, K4, nR2, p->a, p->c, p->gmax, p->alpha, p->d,
                        p->rW, p->W0, Dxd2, K, t, p->dt, p->dx, L);
          t += p->dt;
      } }
    for (int s = 0; s < Sp; s++) for (size_t i = 0; i < n; i++) f[s * n + i] = Rho_t[s][i];
}

/* mean_of_vector / occupied_sites_of_vector (:1211, :1245) */
double ref_riet_mean(const double *x, int n) { vector<double> v(x, x + n); return mean_of_vector(v, n); }
double ref_riet_occupied(const double *x, int n) { vector<double> v(x, x + n); return occupied_sites_of_vector(v, n); }

/* RK4_Wrapper_2D (rietkerk_bjork_basic.cpp:1816): the noise-off runner, r replicates starting from init[r][3][L*L].
 * par = {a, c, gmax, alpha, d, rW, W0, D0, D1, D2, K1, K2, dt, dx, dP}.  rows_out[k][10] = the runner's own output rows
 * {a, r, L, t, <W>, <O>, <P>, var, #surviving, #active} (:2139-2146).  Returns the number of rows, or -1. */
int ref_riet_det_run(const double *par, int L, int r, const double *init, const double *t_meas_in, int n_tmeas,
                     double *rows_out, int rows_cap)
{
    double D[Sp] = {par[7], par[8], par[9]};
    double K[3] = {0.0, par[10], par[11]};
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    D2Vec_Double Rho;
    g_init = init; g_det_replicate = 0; g_frames = nullptr;
    set_global_system_params(par[12], par[13]);
    {
        quiet q;
        RK4_Wrapper_2D(Rho, t_meas, t_meas.back(), par[0], par[1], par[2], par[3], par[4], par[5], par[6], D, K, par[0], par[0],
                       par[12], par[13], par[14], r, L);
    }
    g_init = nullptr;
    const int n = (int)Rho.size();
    if (n > rows_cap) return -1;
    for (int k = 0; k < n; k++)
        for (int j = 0; j < 10; j++) rows_out[10 * k + j] = j < (int)Rho[k].size() ? Rho[k][j] : 0.0;
    return n;
}

}
