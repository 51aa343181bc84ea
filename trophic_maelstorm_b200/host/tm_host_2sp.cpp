// tm_host_2sp.cpp -- see tm_host_2sp.h.  All numerics run in libtm_lattice.so (tm_cons_step2_*); this file is bookkeeping.
#include "tm_host_2sp.h"
#include "tm_host.h"
#include "tm_host_fke.h"
#include "tm_host_side.h"
#include "../../include/tm_fke.h"
#include "../../include/tm_consumer.h"
#include "../../include/tm_lattice.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <sys/stat.h>
#include <iostream>
#include <stdexcept>
#include <string>

namespace tm_b200 { namespace spb2 {

using std::vector;

std::function<void(double *, int, int, double, double, double, const double[])> init_frame;
std::function<void(int, int, double, const double *)> on_measure;
std::string frame_folder, prelim_folder, stat_prefix;
void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern) { init_frame = burnin_frame_rule(parendir, filenamePattern, Sp); }
void set_Prefix(const std::string &user_prefix, double mG, double mP) { set_Prefix_for(SpB, user_prefix, mG, mP, frame_folder, prelim_folder, stat_prefix); }
side::Stepper *test_stepper = nullptr;
const std::string frame_header = "a_c,  x,  P(x; t), G(x; t), W(x; t), O(x; t) \n";            // rietkerk_bjork_constants_2Sp.h:6
const std::string gammaheader = "a_c,  x,  GAM[P(x; t)], GAM[G(x; t)] \n";                      // :8
const std::string movmheader = " a , r, L, t , <GAM[P(x; t)]>_x, Var[<GAM[P(x; t)]>_x], <vx[P(x; t)]>_x,  <vy[P(x; t)]>_x,"
                               " <GAM[G(x; t)]>_x, Var[<GAM[G(x; t)]>_x], <vx[G(x; t)]>_x,  <vy[G(x; t)]>_x, \n";   // :14-15

static bool comparePairs(const std::pair<int, int> &a, const std::pair<int, int> &b)   // R/:108
{
    return (a.first * a.first + a.second * a.second) < (b.first * b.first + b.second * b.second);
}

// R/:1661-1681: the disc of radius R around (0, 0), sorted by squared distance with std::sort (ties keep whatever order
// libstdc++'s introsort leaves them in -- the same library call on the same input as the reference makes)
vector<std::pair<int, int>> computeNeighboringSitesCentral(int R)
{
    vector<std::pair<int, int>> sites;
    for (int x = -R; x <= R; ++x)
        for (int y = -R; y <= R; ++y)
            if (x * x + y * y <= R * R) sites.emplace_back(x, y);
    std::sort(sites.begin(), sites.end(), comparePairs);
    return sites;
}

// R/:1293-1343
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
    tm_cons_ctx *c = nullptr;
    ~Ctx() { if (c) tm_cons_destroy(c); }
    void ck(int rc, const char *what)
    {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tm_cons_last_error(c));
    }
};
// the product's Stepper: the resident SPB = 2 step of tm_consumer.h
struct DeviceStepper : side::Stepper {
    Ctx ctx;
    unsigned long long key = 0;
    size_t n = 0;
    bool stepped = false;
    void upload(const double *frame) override
    {
        ctx.ck(tm_cons_step2_upload(ctx.c, frame, nullptr), "tm_cons_step2_upload");   // gamma starts as zeros (R/:3523)
        stepped = false;
    }
    void run(uint32_t replicate, uint32_t step0, int nsteps) override
    {
        ctx.ck(tm_cons_step2_run(ctx.c, key, replicate, step0, nsteps), "tm_cons_step2_run");
        stepped = true;
    }
    void download(double *frame, double *gamma, double *veff) override
    {
        // the advection vectors are per-replicate state that starts as zeros (R/:3526); the context keeps the last replicate's until a step ran
        ctx.ck(tm_cons_step2_download(ctx.c, frame, gamma, stepped ? veff : nullptr, nullptr, nullptr), "tm_cons_step2_download");
        if (veff && !stepped) std::fill(veff, veff + 2 * n, 0.0);
    }
};
} // namespace

namespace {
// The two replicate runners share everything but the device step: fpe == false: rietkerk_Dornic_2D_MultiSp (Rietkerk_FD, R/:3397);
// fpe == true: rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE, F/:3697), whose step moves the grazer with the FKE kernels instead of a
// Dornic update and clamps at epsilon = 1e-10 (F/:4458).
void run_2sp(bool fpe, D2Vec_Double &Rho, vector<double> &t_meas, double a, double c, double gmax,
             double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB], double (&K)[3],
             double sigma[], double a_st, double a_end, double a_c, double (&A)[SpB][SpB],
             double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[], int (&dtV)[SpB],
             double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar)
{
    if (!init_frame) throw std::runtime_error("tm_b200::spb2::init_frame is not set");
    // dtV[s] (the drivers' round(dtv/dt), R/ order_2stoc_test_rietkerk.cpp): in the two-species runners as shipped it only feeds
    // dtV_counter (R/:3637-3654), which is read by the three-species perception functions (R/:3212, 3324) and by the -DLOCUST_WIND
    // build's direction sampling (R/:3972-3990).  Without that macro the direction is drawn at every step (R/:3992-3995), so any
    // dtV gives the same two-species run: accepted and unused here, as there.
    (void)dtV;
    if (t_meas.empty() || r < 1) throw std::invalid_argument("rietkerk_Dornic_2D_MultiSp: empty t_meas or r < 1");
    const double t_end = t_meas.back();                                               // R/:3439

    const double r_max_effective = (*std::max_element(pR, pR + SpB)) / dx;            // R/:3405
    const double fr_graz = pR[1] / (dx * r_max_effective);                            // R/:3409
    const vector<std::pair<int, int>> nb = computeNeighboringSitesCentral(int(r_max_effective));   // R/:3417
    vector<int32_t> central(2 * nb.size());
    for (size_t k = 0; k < nb.size(); k++) { central[2 * k] = nb[k].first; central[2 * k + 1] = nb[k].second; }

    tm_params veg{};
    veg.model = TM_MODEL_RIETKERK; veg.L = g; veg.dt = dt; veg.dx = dx; veg.a = a;
    veg.rc = c; veg.gmax = gmax; veg.alpha = alpha; veg.rW = rW; veg.W0 = W0; veg.d = M[0]; veg.K1 = K[1]; veg.K2 = K[2];
    veg.D0 = D[0]; veg.D1 = D[2]; veg.D2 = D[3]; veg.sigma0 = sigma[0]; veg.clamp_eps = 1.0e-12; veg.clamp_lo = -1.0e-80;   // R/:4176
    if (fpe) { veg.clamp_eps = fke_globals().epsilon; veg.clamp_lo = 0.0; }                                                 // F/:4458
    tm_graz_params pg{};
    pg.L = g; pg.dt = dt; pg.dx = dx; pg.D = D[1]; pg.v = v[1]; pg.M = M[1]; pg.sigma = sigma[1]; pg.D0_floor = D[0] / (dx * dx);   // R/:4014
    tm_cons_params pc{};
    pc.L = g; pc.dt = dt; pc.a = a; pc.c = c; pc.gmax = gmax; pc.alpha = alpha; pc.rW = rW; pc.W0 = W0; pc.K1 = K[1]; pc.K2 = K[2];
    pc.Dw = D[2] / (dx * dx); pc.Do = D[3] / (dx * dx); pc.A01 = A[0][1]; pc.H01 = H[0][1]; pc.E1 = E[1];               // R/:3463-3480

    DeviceStepper dev;
    side::Stepper *st = test_stepper;     // null in the product: the device context below
    if (!st) {
    Ctx &ctx = dev.ctx;
    dev.key = stream_seed(options().seed, a); dev.n = (size_t)g * g;
    if (tm_cons_create(g, current_device(), &ctx.c) != 0) throw std::runtime_error(std::string("tm_cons_create: ") + tm_cons_last_error(nullptr));
    ctx.ck(tm_cons_step2_setup(ctx.c, &veg, &pg, &pc, central.data(), (int)nb.size(), fr_graz), "tm_cons_step2_setup");
    if (fpe) {
        // the grazer's movement windows and stencils as the runner builds them (F/:3760-3772) from set_global_FKE_params' globals
        const FKEGlobals &G = fke_globals();
        if (G.SpB != SpB || G.L != g) throw std::runtime_error("rietkerk_DorFPE_2D_MultiSp: call set_global_FKE_params(D, v, dtAdv, Dt, dx, g) first");
        const vector<double> gD = precompute_gaussian(sigma[1], G.sig_D_ScaledBounds[1]);
        const vector<double> gV = precompute_gaussian(sigma[1], G.sig_vD_ScaledBounds[1].first);
        tm_fke_terms t{};
        t.bound_D = G.sig_D_ScaledBounds[1];
        t.v_row_lo = t.v_col_lo = G.sig_vD_ScaledBounds[1].first; t.v_row_hi = t.v_col_hi = G.sig_vD_ScaledBounds[1].second;   // CPU path, F/:3640-3641
        t.n_gauss_D = (int)gD.size(); t.n_gauss_V = (int)gV.size();
        t.mu_vel_prefactor = G.mu_vel_prefactor[1]; t.epsilon = G.epsilon;
        ctx.ck(tm_cons_step2_set_fpe(ctx.c, &t, gD.data(), gV.data(), M[1]), "tm_cons_step2_set_fpe");
    }
    st = &dev;
    }

    side::Config cfg;
    cfg.Sp = Sp; cfg.SpB = SpB; cfg.g = g; cfg.fpe = fpe; cfg.a = a; cfg.a_st = a_st; cfg.a_end = a_end; cfg.dt = dt; cfg.dx = dx; cfg.dP = dP;
    cfg.D1 = D[1]; cfg.D2 = D[2]; cfg.Gstar = Gstar; cfg.Vstar = Vstar; cfg.t_max = t_end;
    cfg.frame_folder = frame_folder.empty() ? options().data_root + "/Data/Rietkerk/Frames/Stochastic/2Sp/" + prefix + "_" : frame_folder;
    cfg.prelim_folder = prelim_folder.empty() ? options().data_root + "/Data/Rietkerk/Prelims/Stochastic/2Sp/" + prefix + "_" : prelim_folder;
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
    run_2sp(false, Rho, t_meas, a, c, gmax, alpha, rW, W0, D, v, K, sigma, a_st, a_end, a_c, A, H, E, M, pR, dtV, clow, dt, dx, dP, r, g, Gstar, Vstar);
}

void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax,
                                double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB], double (&K)[3],
                                double sigma[], double a_st, double a_end, double a_c, double (&A)[SpB][SpB],
                                double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[], int (&dtV)[SpB],
                                double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar)
{
    run_2sp(true, Rho, t_meas, a, c, gmax, alpha, rW, W0, D, v, K, sigma, a_st, a_end, a_c, A, H, E, M, pR, dtV, clow, dt, dx, dP, r, g, Gstar, Vstar);
}

const std::string prelimheader = " a , r, L, t , <<W(x; t)>_x>_r, <<O(x; t)>_x>_r,  <<P(x; t)>_x>_r, Var[<P(x; t)>_x]_r, # Surviving Runs P(x; t),"
                                 " # Active Sites P(x; t), <<G(x; t)>_x>_r, Var[<G(x; t)>_x]_r, # Surviving Runs G(x; t), # Active Sites G(x; t), \n";

namespace {
void sweep_2sp(bool fpe, int div, double t_max, double a_start, double a_end, double a_c, double c,
               double gmax, double alpha, double rW, double W0, double (&D)[Sp],
               double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
               double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
               int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
               double Gstar, double Vstar)
{
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = sweep_time_schedule(t_max, dt);                       // R/:4395-4459; sets frame_tmeas
    D2Vec_Double vec;
    vector<D2Vec_Double> per_a(a_space.size());
    for_each_on_devices((int)a_space.size(), [&](int i) {   // the reference's `omp for` over a-values (R/:4509, F/:4795): one GPU per a-value at a time
        vector<double> tm_i = t_measure;
        run_2sp(fpe, per_a[i], tm_i, a_space[i], c, gmax, alpha, rW, W0, D, v, K, sigma, a_start, a_end, a_c, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
    });
    for (const D2Vec_Double &rows : per_a) vec.insert(vec.end(), rows.begin(), rows.end());
    std::stringstream L, tm, d3, p1, p2, rini, Dm1, dix, geq, veq;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << a_start; p2 << a_end; rini << r;
    Dm1 << std::setprecision(3) << D[1]; dix << std::setprecision(2) << dx; geq << std::setprecision(5) << Gstar; veq << std::setprecision(5) << Vstar;
    // Rietkerk_FD's header default: "../Data/Rietkerk/Stochastic/2Sp/1stCC_Rietkerk_<prefix>_P_c_G_"; Rietkerk_FPE's set_Prefix (F/:143-165):
    // "../Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_"
    const std::string dir = options().data_root + (fpe ? "/Data/Rietkerk/Stochastic" : "/Data/Rietkerk/Stochastic/2Sp");
    std::string savefile = (stat_prefix.empty() ? dir + (fpe ? "/1stOrderCC_Rietkerk_" : "/1stCC_Rietkerk_") + prefix + "_P_c_G_" : stat_prefix) + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_D1_" + Dm1.str()
                         + "_a1_" + p1.str() + "_a2_" + p2.str() + "_dx_" + dix.str();                                    // R/:4582, F/:4852
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
    sweep_2sp(false, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
}

void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                           double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                           int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar, double Vstar)
{
    sweep_2sp(true, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E, M, pR, dtV, clo, dt, dx, dP, r, g, Gstar, Vstar);
}

}} // namespace tm_b200::spb2
