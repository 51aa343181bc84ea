/*
 * oracle/ref_harness_fke_files.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * The UNMODIFIED reference runner rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE/rietkerk_bjork_FKE.cpp:3697) compiled with -DSPB=2 (CPU
 * movement path) WITH ITS OWN CSV WRITERS (save_frame :2829, save_prelimframe :2723, the wrappers :2879 / :3024): only the frame
 * initialiser the runner calls as shipped (init_gaussframe :3962) is replaced.  It writes its whole output tree under a directory of the
 * caller's choice; tests/golden/make_side_outputs_golden.py turns the tree into the fixture the host layer's writers are compared with
 * (the Rietkerk_FPE flavour of the bookkeeping: checkpoints at the 85 / 90 / 95 / 100 % bins only, its own frame-saving rule).
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

extern "C" {

/* par as ref_fke_run (oracle/ref_harness_fke.cpp: the SPB = 2 runner vector plus par[23] = dtAdv[1]); r replicates, replicate j starts
 * from init frame j % n_init; root stands for the reference's "../Data/Rietkerk/" (must exist); user_prefix: set_Prefix's argument.
 * Returns the uniforms consumed. */
uint64_t reffkef_run(const double *par, int L, int r, const double *init, int n_init, const double *t_meas_in, int n_tmeas,
                     const double *frame_t_in, int n_frame_t, const double *u, uint64_t n_u, const char *root, const char *user_prefix,
                     double Gstar, uint64_t *overflow)
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
    string pre = user_prefix;
    set_Prefix(pre, 0.0, 0.0);
    const string rt = root;                       /* "../Data/Rietkerk/" -> root */
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
