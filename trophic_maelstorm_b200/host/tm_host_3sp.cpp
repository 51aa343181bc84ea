// tm_host_3sp.cpp -- see tm_host_3sp.h.  Host bookkeeping only; the lattice numerics go through include/tm_consumer3.h.
#include "tm_host_3sp.h"
#include "tm_host.h"
#include "tm_host_2sp.h"   // computeNeighboringSitesCentral (R/:1661, the same for every SPB)
#include "tm_host_fke.h"   // fke_globals, precompute_gaussian
#include "tm_host_side.h"
#include "../../include/tm_fke.h"
#include "../../include/tm_consumer3.h"
#include "../../include/tm_lattice.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>

namespace tm_b200 { namespace spb3 {

using std::vector;

std::function<void(double *, int, int, double, double, double, const double[])> init_frame;
std::function<void(int, int, double, const double *)> on_measure;
std::string frame_folder, prelim_folder, stat_prefix;
void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern) { init_frame = burnin_frame_rule(parendir, filenamePattern, Sp); }
void set_Prefix(const std::string &user_prefix, double mG, double mP) { set_Prefix_for(SpB, user_prefix, mG, mP, frame_folder, prelim_folder, stat_prefix); }
side::Stepper *test_stepper = nullptr;
const std::string frame_header = "a_c,  x,  P(x; t), G(x; t), Pr(x; t), W(x; t), O(x; t) \n";            // rietkerk_bjork_constants_3Sp.h:6
const std::string gammaheader = "a_c,  x,  GAM[P(x; t)], GAM[G(x; t)], GAM[Pr(x; t)] \n";                 // :8
const std::string movmheader = " a , r, L, t , <GAM[P(x; t)]>_x, Var[<GAM[P(x; t)]>_x], <vx[P(x; t)]>_x,  <vy[P(x; t)]>_x,"
                               " <GAM[G(x; t)]>_x, Var[<GAM[G(x; t)]>_x], <vx[G(x; t)]>_x,  <vy[G(x; t)]>_x,"
                               " <GAM[Pr(x; t)]>_x, Var[<GAM[Pr(x; t)]>_x], <vx[Pr(x; t)]>_x,  <vy[Pr(x; t)]>_x, \n";   // :15-17

// R/:1293-1343 with SpB = 3 species blocks and the two water columns
void var_mean_incremental_surv_runs(D2Vec_Double &T, const D2Vec_Double &X, int size, int j)
{
    vector<double> mean_prev(size);
    for (int s = 0; s < SpB; s++)
        for (int i = 0; i < size; i++) {
            const int r = (int)(T[i][4 * s + 5] + 1);
            if (X[i][2 * s + 0] == 0.0) break;
            mean_prev[i] = T[i][4 * s + 3];
            if (T[i][4 * s + 5] == 0.0) {
                T[i][4 * s + 3] = X[i][2 * s + 1]; T[i][4 * s + 4] = 0.0; T[i][4 * s + 6] = X[i][2 * s + 0];
            } else {
                T[i][4 * s + 3] = (X[i][2 * s + 1] + (r - 1) * T[i][4 * s + 3]) / r;
                T[i][4 * s + 6] = (X[i][2 * s + 0] + (r - 1) * T[i][4 * s + 6]) / r;
                T[i][4 * s + 4] = (T[i][4 * s + 3] - mean_prev[i]) * (T[i][4 * s + 3] - mean_prev[i])
                                + (1 / (r - 1)) * ((r - 2) * T[i][4 * s + 4]
                                                   + (X[i][2 * s + 1] - T[i][4 * s + 3]) * (X[i][2 * s + 1] - T[i][4 * s + 3]));   // int 1/(r-1)
            }
            T[i][4 * s + 5] = r;
        }
    for (int i = 0; i < size; i++) {
        T[i][1] = (X[i][2 * (Sp - 2) + 1] + (j)*T[i][1]) / (j + 1);
        T[i][2] = (X[i][2 * (Sp - 1) + 1] + (j)*T[i][2]) / (j + 1);
    }
}

namespace {
struct Ctx {
    tm_cons3_ctx *c = nullptr;
    ~Ctx() { if (c) tm_cons3_destroy(c); }
    void ck(int rc, const char *what)
    {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tm_cons3_last_error(c));
    }
};
// the product's Stepper: the resident SPB = 3 step of tm_consumer3.h
struct DeviceStepper : side::Stepper {
    Ctx ctx;
    unsigned long long key = 0;
    size_t n = 0;
    bool stepped = false;
    void upload(const double *frame) override
    {
        ctx.ck(tm_cons3_upload(ctx.c, frame, nullptr), "tm_cons3_upload");             // gamma starts as zeros (R/:3523)
        stepped = false;
    }
    void run(uint32_t replicate, uint32_t step0, int nsteps) override
    {
        ctx.ck(tm_cons3_run(ctx.c, key, replicate, step0, nsteps), "tm_cons3_run");
        stepped = true;
    }
    void download(double *frame, double *gamma, double *veff) override
    {
        // the advection vectors are per-replicate state that starts as zeros (R/:3526); the context keeps the last replicate's until a step ran
        ctx.ck(tm_cons3_download(ctx.c, frame, gamma, stepped ? veff : nullptr, nullptr, nullptr), "tm_cons3_download");
        if (veff && !stepped) std::fill(veff, veff + 4 * n, 0.0);
    }
};
} // namespace

namespace {
// fpe == false: rietkerk_Dornic_2D_MultiSp (Rietkerk_FD, R/:3397); true: rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE, F/:3697)
void run_3sp(bool fpe, D2Vec_Double &Rho, vector<double> &t_meas, double a, double c, double gmax,
             double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB], double (&K)[3],
             double sigma[], double a_st, double a_end, double a_c, double (&A)[SpB][SpB],
             double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[], int (&dtV)[SpB],
             double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar)
{
    if (!init_frame) throw std::runtime_error("tm_b200::spb3::init_frame is not set");
    (void)dtV;   // see the header
    if (t_meas.empty() || r < 1) throw std::invalid_argument("rietkerk_Dornic_2D_MultiSp: empty t_meas or r < 1");
    const double t_end = t_meas.back();                                               // R/:3439

    const double r_max_effective = (*std::max_element(pR, pR + SpB)) / dx;            // R/:3405
    const double fr_graz = pR[1] / (dx * r_max_effective), fr_pred = pR[2] / (dx * r_max_effective);   // R/:3409
    const vector<std::pair<int, int>> nb = spb2::computeNeighboringSitesCentral(int(r_max_effective));   // R/:3417
    vector<int32_t> central(2 * nb.size());
    for (size_t k = 0; k < nb.size(); k++) { central[2 * k] = nb[k].first; central[2 * k + 1] = nb[k].second; }

    tm_params veg{};
    veg.model = TM_MODEL_RIETKERK; veg.L = g; veg.dt = dt; veg.dx = dx; veg.a = a;
    veg.rc = c; veg.gmax = gmax; veg.alpha = alpha; veg.rW = rW; veg.W0 = W0; veg.d = M[0]; veg.K1 = K[1]; veg.K2 = K[2];
    veg.D0 = D[0]; veg.D1 = D[3]; veg.D2 = D[4]; veg.sigma0 = sigma[0]; veg.clamp_eps = 1.0e-12; veg.clamp_lo = -1.0e-80;   // R/:4176
    if (fpe) { veg.clamp_eps = fke_globals().epsilon; veg.clamp_lo = 0.0; }                                                 // F/:4458
    tm_graz_params pg[2];
    for (int s = 1; s <= 2; s++) {
        tm_graz_params &q = pg[s - 1];
        q = tm_graz_params{};
        q.L = g; q.dt = dt; q.dx = dx; q.D = D[s]; q.v = v[s]; q.M = M[s]; q.sigma = sigma[s]; q.D0_floor = D[0] / (dx * dx);   // R/:4014
    }
    tm_cons3_params pc{};
    pc.L = g; pc.dt = dt; pc.a = a; pc.c = c; pc.gmax = gmax; pc.alpha = alpha; pc.rW = rW; pc.W0 = W0; pc.K1 = K[1]; pc.K2 = K[2];
    pc.Dw = D[3] / (dx * dx); pc.Do = D[4] / (dx * dx);                               // R/:3463-3480
    pc.A01 = A[0][1]; pc.H01 = H[0][1]; pc.E1 = E[1]; pc.A12 = A[1][2]; pc.H12 = H[1][2]; pc.E2 = E[2];

    DeviceStepper dev;
    side::Stepper *st = test_stepper;     // null in the product: the device context below
    if (!st) {
    Ctx &ctx = dev.ctx;
    dev.key = stream_seed(options().seed, a); dev.n = (size_t)g * g;
    if (tm_cons3_create(g, current_device(), &ctx.c) != 0) throw std::runtime_error(std::string("tm_cons3_create: ") + tm_cons3_last_error(nullptr));
    ctx.ck(tm_cons3_setup(ctx.c, &veg, &pg[0], &pg[1], &pc, central.data(), (int)nb.size(), fr_graz, fr_pred), "tm_cons3_setup");
    if (fpe) {
        // movement windows and stencils of grazer and predator as the runner builds them (F/:3760-3772) from set_global_FKE_params' globals
        const FKEGlobals &G = fke_globals();
        if (G.SpB != SpB || G.L != g) throw std::runtime_error("rietkerk_DorFPE_2D_MultiSp: call set_global_FKE_params(D, v, dtAdv, Dt, dx, g) first");
        vector<double> gD[2], gV[2];
        tm_fke_terms t[2];
        const double *pD[2], *pV[2];
        for (int s = 1; s <= 2; s++) {
            gD[s - 1] = precompute_gaussian(sigma[s], G.sig_D_ScaledBounds[s]);
            gV[s - 1] = precompute_gaussian(sigma[s], G.sig_vD_ScaledBounds[s].first);
            tm_fke_terms &q = t[s - 1];
            q = tm_fke_terms{};
            q.bound_D = G.sig_D_ScaledBounds[s];
            q.v_row_lo = q.v_col_lo = G.sig_vD_ScaledBounds[s].first; q.v_row_hi = q.v_col_hi = G.sig_vD_ScaledBounds[s].second;   // CPU path, F/:3640-3641
            q.n_gauss_D = (int)gD[s - 1].size(); q.n_gauss_V = (int)gV[s - 1].size();
            q.mu_vel_prefactor = G.mu_vel_prefactor[s]; q.epsilon = G.epsilon;
            pD[s - 1] = gD[s - 1].data(); pV[s - 1] = gV[s - 1].data();
        }
        ctx.ck(tm_cons3_set_fpe(ctx.c, t, pD, pV, M[1], M[2]), "tm_cons3_set_fpe");
    }
    st = &dev;
    }

    side::Config cfg;
    cfg.Sp = Sp; cfg.SpB = SpB; cfg.g = g; cfg.fpe = fpe; cfg.a = a; cfg.a_st = a_st; cfg.a_end = a_end; cfg.dt = dt; cfg.dx = dx; cfg.dP = dP;
    cfg.D1 = D[1]; cfg.D2 = D[2]; cfg.Gstar = Gstar; cfg.Vstar = Vstar; cfg.t_max = t_end;
    cfg.frame_folder = frame_folder.empty() ? options().data_root + "/Data/Rietkerk/Frames/Stochastic/3Sp/" + prefix + "_" : frame_folder;
    cfg.prelim_folder = prelim_folder.empty() ? options().data_root + "/Data/Rietkerk/Prelims/Stochastic/3Sp/" + prefix + "_" : prelim_folder;
    cfg.frame_header = frame_header; cfg.gammaheader = gammaheader; cfg.prelimheader = prelimheader; cfg.movmheader = movmheader;
    side::Run run;
    run.r = r; run.a_c = a_c; run.clow = clow; run.init_frame = init_frame; run.on_measure = on_measure;
    run.var_mean = [](D2Vec_Double &T, const D2Vec_Double &X, int size, int j) { var_mean_incremental_surv_runs(T, X, size, j); };
    side::run_replicates(*st, cfg, run, t_meas, Rho);
}

} // namespace

void rietkerk_Dornic_2D_MultiSp(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax,
                                double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB], double (&K)[3],
                                double sigma[], double a_st, double a_end, double a_c, double (&A)[SpB][SpB],
                                double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[], int (&dtV)[SpB],
                                double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar)
{
    run_3sp(false, Rho, t_meas, a, c, gmax, alpha, rW, W0, D, v, K, sigma, a_st, a_end, a_c, A, H, E, M, pR, dtV, clow, dt, dx, dP, r, g, Gstar, Vstar);
}

void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax,
                                double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB], double (&K)[3],
                                double sigma[], double a_st, double a_end, double a_c, double (&A)[SpB][SpB],
                                double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[], int (&dtV)[SpB],
                                double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar)
{
    run_3sp(true, Rho, t_meas, a, c, gmax, alpha, rW, W0, D, v, K, sigma, a_st, a_end, a_c, A, H, E, M, pR, dtV, clow, dt, dx, dP, r, g, Gstar, Vstar);
}

const std::string prelimheader = " a , r, L, t , <<W(x; t)>_x>_r, <<O(x; t)>_x>_r,  <<P(x; t)>_x>_r, Var[<P(x; t)>_x]_r, # Surviving Runs P(x; t),"
                                 " # Active Sites P(x; t), <<G(x; t)>_x>_r, Var[<G(x; t)>_x]_r, # Surviving Runs G(x; t), # Active Sites G(x; t),"
                                 "<<Pr(x; t)>_x>_r, Var[<Pr(x; t)>_x]_r, # Surviving Runs Pr(x; t), # Active Sites Pr(x; t), \n";   // constants_3Sp.h:10-12

namespace {
void sweep_3sp(bool fpe, int div, double t_max, double a_start, double a_end, double a_c, double c,
               double gmax, double alpha, double rW, double W0, double (&D)[Sp],
               double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
               double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
               int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
               double Gstar, double Vstar)
{
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = sweep_time_schedule(t_max, dt);                       // R/:4395-4459; sets frame_tmeas
    vector<D2Vec_Double> per_a(a_space.size());
    for_each_on_devices((int)a_space.size(), [&](int i) {   // the reference's `omp for` over a-values (R/:4509): one GPU per a-value at a time
        vector<double> tm_i = t_measure;
        run_3sp(fpe, per_a[i], tm_i, a_space[i], c, gmax, alpha, rW, W0, D, v, K, sigma, a_start, a_end, a_c, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
    });
    D2Vec_Double vec;
    for (const D2Vec_Double &rows : per_a) vec.insert(vec.end(), rows.begin(), rows.end());
    std::stringstream L, tm, d3, p1, p2, rini, Dm1, dix, geq, veq;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << a_start; p2 << a_end; rini << r;
    Dm1 << std::setprecision(3) << D[1]; dix << std::setprecision(2) << dx; geq << std::setprecision(5) << Gstar; veq << std::setprecision(5) << Vstar;
    // Rietkerk_FD: "../Data/Rietkerk/Stochastic/3Sp/1stCC_Rietkerk_<prefix>_P_c_G_"; Rietkerk_FPE's set_Prefix (F/:143-165):
    // "../Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_"
    const std::string dir = options().data_root + (fpe ? "/Data/Rietkerk/Stochastic" : "/Data/Rietkerk/Stochastic/3Sp");
    std::string savefile = (stat_prefix.empty() ? dir + (fpe ? "/1stOrderCC_Rietkerk_" : "/1stCC_Rietkerk_") + prefix + "_P_c_G_" : stat_prefix) + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_D1_" + Dm1.str()
                         + "_a1_" + p1.str() + "_a2_" + p2.str() + "_dx_" + dix.str();                                    // R/:4582
    if (Gstar != -1) savefile += "_Geq_" + geq.str() + "_R_" + rini.str() + ".csv";
    else if (Vstar != -1) savefile += "_Veq_" + veq.str() + "_R_" + rini.str() + ".csv";
    else savefile += "_R_" + rini.str() + ".csv";
    recursive_dir_create(savefile.substr(0, savefile.find_last_of('/')));
    std::ofstream out(savefile);
    out << prelimheader << "\n";                                                     // R/:4601
    for (size_t i = 0; i < vec.size(); i++) {
        for (size_t j = 0; j + 1 < vec[i].size(); j++) out << std::setprecision(16) << vec[i][j] << ",";
        out << std::setprecision(16) << vec[i][vec[i].size() - 1] << std::endl;
    }
}
} // namespace

void first_order_critical_exp_delta_stochastic_MultiSp(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                       double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                       double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                       double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                       int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                       double Gstar, double Vstar)
{
    sweep_3sp(false, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
}

void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                           double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                           int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar, double Vstar)
{
    sweep_3sp(true, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
}

}} // namespace tm_b200::spb3
