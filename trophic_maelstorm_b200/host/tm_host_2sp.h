// tm_host_2sp.h -- reference-facing host layer for the vegetation + grazer model (SPB = 2, Sp = 4 fields P, G, W, O):
// the reference's runner signature over the resident device stepper of include/tm_consumer.h.
//
//   rietkerk_Dornic_2D_MultiSp       Rietkerk_FD/rietkerk_bjork_basic.cpp:3397 compiled with -DSPB=2
//   computeNeighboringSitesCentral   Rietkerk_FD/rietkerk_bjork_basic.cpp:1661
//   var_mean_incremental_surv_runs   Rietkerk_FD/rietkerk_bjork_basic.cpp:1293 (SpB species blocks + the two water columns)
//
// The reference compiles one binary per SPB with the same function names; here the SPB = 2 set lives in its own
// namespace (tm_b200::spb2) so that both can be linked into one program.  Kept from the reference: the measurement
// schedule from the accumulated t += dt (R/:3648, 3659), Rho_M = {#occupied, <rho>} per field from left-to-right sums of
// the downloaded frame (mean_of_vector R/:1211), the stop of a replicate whose vegetation is gone at a checkpoint bin
// (R/:3710-3716), the replicate-order running mean/variance with the integer 1/(r-1) (R/:1325), and the row layout
// {a, r, L, t, <W>, <O>, then per living species <rho>, var, #surviving, #active} (R/:4361-4372), and the
// runner's CSV side outputs (tm_host_side.h): FRAME_* / GAMMA_* snapshots on tm_b200::frame_tmeas, the movement averages Rho_Mov
// and the PRELIM_TSERIES_* / PRELIM_MOVSERIES_* checkpoints and PRELIM_AGGRAND_* running tables while tm_b200::write_prelims is on.
// dtV[s] is accepted and has no effect, as in the reference's two-species build (only its -DLOCUST_WIND variant and the three-species
// perception functions read the counter it feeds).
// The Rietkerk_FPE runner of the same model (rietkerk_DorFPE_2D_MultiSp) is here too: same bookkeeping, different device step.
#pragma once
#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace tm_b200 { namespace side { struct Stepper; }
namespace spb2 {

// TEST SEAM.  Null in the product (the runners then step on the device, and fail without one).  When set, the runners hand their
// bookkeeping this Stepper instead of creating a device context: the CPU tests pin the measurement schedule, the running statistics
// and the CSV writers on the reference's own output tree this way.
extern side::Stepper *test_stepper;

using D2Vec_Double = std::vector<std::vector<double>>;
constexpr int Sp = 4, SpB = 2;

// Initial frame of replicate j: 4*g*g doubles (P, G, W, O).  The reference builds it from exprtk expression strings
// (init_exprtk_randbiMFTframe R/:396); here the caller supplies the rule.  Must be set before the runner is called.
extern std::function<void(double *frame, int g, int replicate, double a, double a_c, double dP, const double clow[])> init_frame;
// Optional observer called at every measurement bin with the downloaded frame (4*g*g) -- what the reference's frame
// writers see; used by the tests.
extern std::function<void(int replicate, int index, double t, const double *frame)> on_measure;
// Output locations of the side outputs; empty (default) = "<data_root>/Data/Rietkerk/{Frames,Prelims}/Stochastic/2Sp/<prefix>_"
// with tm_b200::prefix (the reference's frame_folder / prelim_folder, header :106-107).
extern std::string frame_folder, prelim_folder;
// set_Prefix (R/:113-135, F/:143-165) of the -DSPB=2 binaries: tm_b200::prefix and the three output locations
// "<data_root>/Data/Rietkerk/{Frames,Prelims}/Stochastic/2Sp[_mG_<mG>[_mP_<mP>]]/<prefix>_", ".../Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_"
extern std::string stat_prefix;   // empty (default): the header's location of the final table, ".../Stochastic/2Sp/1stCC_Rietkerk_<prefix>_P_c_G_"
void set_Prefix(const std::string &user_prefix, double mG = 0.0, double mP = 0.0);
// burn-in frames (the reference's -DINIT=2 path, init_burnin_wrapper R/:836): every replicate starts from "<parendir><pattern><r'>.csv", a
// frame in save_frame()'s layout with this model's columns (located by header name, water last)
void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern);
extern const std::string frame_header, gammaheader, movmheader;   // rietkerk_bjork_constants_2Sp.h:6-15

std::vector<std::pair<int, int>> computeNeighboringSitesCentral(int R);
void var_mean_incremental_surv_runs(D2Vec_Double &t_avg_var_rep_N, const D2Vec_Double &X_curr, int size, int j);

void rietkerk_Dornic_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);

// The Rietkerk_FPE twin, Rietkerk_FPE/rietkerk_bjork_FKE.cpp:3697 compiled with -DSPB=2: same signature, same bookkeeping; its device
// step (tm_cons_step2_set_fpe) has no grazer Dornic update, carries the grazer's mortality in the RK4 reaction terms, moves the
// grazer with the FKE kernels on the device-resident field (the reference's -DBARRACUDA seam copies three fields to the device and
// one back every step, F/:4282-4299) and clamps at epsilon = 1e-10.  Call tm_b200::set_global_FKE_params (tm_host_fke.h) first, as
// the reference's drivers do (order_2stoc_unity_rietkerk.cpp:202): the movement windows and stencils come from its globals.
void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);
// its a-sweep (first_order_critical_exp_delta_stochastic_MultiSp of Rietkerk_FPE, F/:4665): CSV
// "<data_root>/Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_<prefix>_P_c_G_<L>_T_..._R_<r>.csv" (set_Prefix F/:143-165)
void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                           double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                           int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar = -1.0, double Vstar = -1.0);

// The a-sweep driver entry, R/:4379-4617 compiled with -DSPB=2 (a-values dealt over the visible GPUs, tm_host.h): runner per a-value, one CSV
// "<data_root>/Data/Rietkerk/Stochastic/2Sp/1stCC_Rietkerk_<prefix>_P_c_G_<L>_T_<T>_dt_..._R_<r>.csv" with the reference's
// 2Sp header (rietkerk_bjork_constants_2Sp.h:10-11) and 16-digit rows.
extern const std::string prelimheader;
void first_order_critical_exp_delta_stochastic_MultiSp(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                       double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                       double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                       double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                       int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                       double Gstar = -1.0, double Vstar = -1.0);

}} // namespace tm_b200::spb2
