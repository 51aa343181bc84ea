// tm_host_3sp.h -- reference-facing host layer for the vegetation + grazer + predator model (SPB = 3, Sp = 5 fields P, G, Pr, W, O): the
// reference's runner signature over the resident device stepper of include/tm_consumer3.h.
//
//   rietkerk_Dornic_2D_MultiSp                           Rietkerk_FD/rietkerk_bjork_basic.cpp:3397 compiled with -DSPB=3
//   first_order_critical_exp_delta_stochastic_MultiSp    :4379 (its a-sweep)
//
// The reference compiles one binary per SPB with the same function names; here each set lives in its own namespace (tm_b200 for
// SPB = 1, tm_b200::spb2, tm_b200::spb3).  Kept from the reference, as in the SPB = 2 layer: the measurement schedule from the
// accumulated t += dt (R/:3648, 3659), Rho_M = {#occupied, <rho>} per field from left-to-right sums of the downloaded frame, the stop
// of a replicate whose vegetation is gone at a checkpoint bin (R/:3710-3716), the replicate-order running mean/variance with the
// integer 1/(r-1) (R/:1325) and the row layout {a, r, L, t, <W>, <O>, then per living species <rho>, var, #surviving, #active}
// (R/:4361-4372: 18 columns).  dtV is accepted and unused (the counter it feeds is read only inside commented-out code of
// calc_gamma_3Sp_NonRefugia, R/:3211-3214, and by the -DLOCUST_WIND build).  The runner's CSV side outputs (FRAME_* / GAMMA_* snapshots,
// PRELIM_TSERIES_* / PRELIM_MOVSERIES_* checkpoints, PRELIM_AGGRAND_* running tables) are written as in the SPB = 2 layer (tm_host_side.h).
#pragma once
#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace tm_b200 { namespace side { struct Stepper; }
namespace spb3 {

extern side::Stepper *test_stepper;   // TEST SEAM, null in the product: see tm_host_2sp.h

using D2Vec_Double = std::vector<std::vector<double>>;
constexpr int Sp = 5, SpB = 3;

// Initial frame of replicate j: 5*g*g doubles (P, G, Pr, W, O).  Must be set before the runner is called.
extern std::function<void(double *frame, int g, int replicate, double a, double a_c, double dP, const double clow[])> init_frame;
// Optional observer called at every measurement bin with the downloaded frame (5*g*g).
extern std::function<void(int replicate, int index, double t, const double *frame)> on_measure;
// Output locations of the side outputs; empty (default) = "<data_root>/Data/Rietkerk/{Frames,Prelims}/Stochastic/3Sp/<prefix>_"
extern std::string frame_folder, prelim_folder;
// set_Prefix (R/:113-135, F/:143-165) of the -DSPB=3 binaries: tm_b200::prefix and the three output locations
// "<data_root>/Data/Rietkerk/{Frames,Prelims}/Stochastic/3Sp[_mG_<mG>[_mP_<mP>]]/<prefix>_", ".../Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_"
extern std::string stat_prefix;   // empty (default): the header's location of the final table, ".../Stochastic/3Sp/1stCC_Rietkerk_<prefix>_P_c_G_"
void set_Prefix(const std::string &user_prefix, double mG = 0.0, double mP = 0.0);
// burn-in frames (the reference's -DINIT=2 path, init_burnin_wrapper R/:836): every replicate starts from "<parendir><pattern><r'>.csv", a
// frame in save_frame()'s layout with this model's columns (located by header name, water last)
void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern);
extern const std::string frame_header, gammaheader, movmheader;   // rietkerk_bjork_constants_3Sp.h:6-17

void var_mean_incremental_surv_runs(D2Vec_Double &t_avg_var_rep_N, const D2Vec_Double &X_curr, int size, int j);   // R/:1293

void rietkerk_Dornic_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);

// The Rietkerk_FPE twin, Rietkerk_FPE/rietkerk_bjork_FKE.cpp:3697 compiled with -DSPB=3 (device step: tm_cons3_set_fpe): call
// tm_b200::set_global_FKE_params (tm_host_fke.h) first, as the reference's drivers do (order_3stoc_unity_rietkerk.cpp).
void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);
// its a-sweep (F/:4665): CSV "<data_root>/Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_<L>_T_..._R_<r>.csv"
void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                           double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                           int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar = -1.0, double Vstar = -1.0);

// CSV "<data_root>/Data/Rietkerk/Stochastic/3Sp/1stCC_Rietkerk_<prefix>_P_c_G_<L>_T_<T>_dt_..._R_<r>.csv" with the reference's 3Sp header
// (rietkerk_bjork_constants_3Sp.h:10-12) and 16-digit rows; a-values dealt over the visible GPUs (tm_host.h).
extern const std::string prelimheader;
void first_order_critical_exp_delta_stochastic_MultiSp(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                       double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                       double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                       double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                       int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                       double Gstar = -1.0, double Vstar = -1.0);

}} // namespace tm_b200::spb3
