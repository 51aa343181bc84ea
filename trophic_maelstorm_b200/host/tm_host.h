// tm_host.h -- the reference-facing host layer: the reference's own C++ entry points for the
// Dornic path, same names and argument lists, implemented over the C ABI (include/tm_lattice.h).
// A reference driver main switches over by including this header, `using namespace tm_b200;`
// and linking libtm_host.so + libtm_lattice.so.
//
//   percolationDornic_2D            Percolation_FD/stochasticSPDP.cpp:940
//   expanded_percolationDornic_2D   Percolation_FD/stochasticSPDP.cpp:1809
//   critical_exp_delta              Percolation_FD/stochasticSPDP.cpp:1309
//   first_order_critical_exp_delta  Percolation_FD/stochasticSPDP.cpp:2422
//   critical_exp_finite_scaling_collapse / _stationary   Percolation_FD/stochasticSPDP.cpp:1424, 1538
//   terminal_percolationDornic_2D, first_order_terminal_diff   Percolation_FD/stochasticSPDP.cpp:2158, 2548
//   frames_expanded_percolationDornic_2D, capture_frame_decay  Percolation_FD/stochasticSPDP.cpp:2678, 2997
//   rietkerk_Dornic_2D_MultiSp      Rietkerk_FD/rietkerk_bjork_basic.cpp:3397   (SPB = 1)
//   first_order_critical_exp_delta_stochastic_MultiSp   Rietkerk_FD/rietkerk_bjork_basic.cpp:4379
//   spreading_exp_Dornic_2D         Percolation_FD/stochasticSPDP.cpp:1145   (seed-at-centre runs: N(t), R^2(t), P_surv(t))
//   critical_exp_theta_delta        Percolation_FD/stochasticSPDP.cpp:1667
//   logarithm10_time_bins, linspace, var_mean_incremental_surv_runs, spreading_incremental_surv_runs,
//   init_solitarytear, save_frame  (bookkeeping)
//
// What is different by construction: replicates run concurrently as trials of one device context
// instead of a serial `for j<r` loop, and random numbers come from Philox streams keyed by
// (seed, trial, step, cell) instead of std::mt19937_64 seeded from std::random_device.
#pragma once
#include <functional>
#include <string>
#include <vector>

namespace tm_b200 {

typedef std::vector<std::vector<double>> D2Vec_Double;

struct Options {
    int device = 0;
    unsigned long long seed = 0x5EEDC0DEull; // the reference seeds from std::random_device (P/:973); here it is explicit
    int n_devices = 0;                        // GPUs device .. device + n_devices - 1 share a sweep (0: every visible GPU from `device` on)
    int max_trials_per_ctx = 0;               // cap on the trials one device context holds (0: from the free device memory)
    std::vector<int> device_ids;              // explicit worker list (overrides device / n_devices); an id may repeat: that GPU then
                                              // serves two workers at once (env TM_DEVICE_LIST=0,1,1)
    bool strict = false;                      // reference operation order in the RK4 stages
    bool quiet = true;                        // the reference's progress chatter on stdout
    std::string data_root = "..";             // outputs go to <data_root>/Data/... like the reference's "../Data/..."
    bool file_ids = false;                    // the two random tokens of the Rietkerk file names ("jID_<n>": thrID R/:3634, an integer in
                                              // 0..10000 per a-value; "_ID_<x>_": R/:4343, a real in [0,1) per replicate) drawn from
                                              // (seed, a, replicate) instead of the constant 0 (env TM_FILE_IDS=1)
};
Options &options(); // process-wide; env TM_DEVICE, TM_SEED, TM_STRICT, TM_DATA_ROOT, TM_FILE_IDS override the defaults
// What a driver main calls first.  The reference seeds every replicate from std::random_device: two runs of a program are independent and
// -- through the random tokens in their file names -- leave their frames and tables side by side in the same folders.  Without TM_SEED in
// the environment this does the same: options().seed from std::random_device (printed, so that the run can be repeated with TM_SEED=...)
// and file_ids on.  With TM_SEED set the run is reproducible: that seed, and the constant tokens unless TM_FILE_IDS=1.
void independent_run_defaults();
std::string jid_token(double a);              // "0", or the a-value's thrID when options().file_ids
std::string id_token(double a, int replicate); // "0", or the replicate's ID

// The Philox key a runner hands to the device for control parameter `a`: options().seed mixed with the bit pattern of a.
// The reference seeds every replicate of every a-value from std::random_device (P/:973, R/:3511), so the a-values of a
// sweep are statistically independent; a fixed key would give trials 0..r-1 the same (seed, trial, step, cell) counters
// at every a-value.  The mix depends on (seed, a) only -- not on call order or on which GPU runs the a-value -- so a
// sweep is reproducible and identical on 1 or 8 GPUs.  Runs meant to be independent need different TM_SEED values.
unsigned long long stream_seed(unsigned long long seed, double a);

// The ensemble scheduler (the reference's `#pragma omp parallel for` over a-values, R/:4509-4542, P/:2462-2488): runs
// fn(0) .. fn(n_items - 1), item i on GPU device_list()[i mod #GPUs], one host thread per GPU, items of a GPU in
// increasing order.  Inside fn every tm_create of this layer binds current_device().
std::vector<int> device_list();
int current_device();
void for_each_on_devices(int n_items, const std::function<void(int)> &fn);

// ---- bookkeeping (host) --------------------------------------------------------------------
std::vector<double> linspace(double start, double end, int num);                 // P/:254
std::vector<double> logarithm10_time_bins(double t_max, double dt);              // P/:336
void var_mean_incremental_surv_runs(double t_avg_var_rep_N[][5], double X_curr[][2], int size); // P/:173

// ---- DP --------------------------------------------------------------------------------------
void percolationDornic_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_meas, double Rh0[],
                          double t_max, double a, double b, double D, double sigma, double dt, double dx, int r, int g);
// noise-off runner, P/:549-662: classical RK4 over fP_2D (fourth-order Laplacian), ONE replicate whatever r is (the reference's
// loop is `for (j = 0; j < 1; j++)`); rows {a, replicate, t, cell, rho(cell)} for every cell at every t_stop window
void percNoDiff_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_stop, double Rh0[], double t_max, double a,
                   double b, double D, double dt, double dx, int r, int g);
void expanded_percolationDornic_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_meas, double Rh0[],
                                   double t_max, double a, double b, double c, double D, double sigma, double dt,
                                   double dx, int r, int g);
void critical_exp_delta(double Rho_0[], int div, double t_max, double a_start, double a_end, double b, double D,
                        double sigma, double dt, double dx, int r, int g);
// terminal state after a fixed time (P/:2158, :2548; driver bifurc_prob.cpp): the cubic DP step from a given frame, r replicates, rows
// {a, c, L, j, t, <rho(0)>, <rho(t)>, #active(t)}; the sweep over constant initial densities p writes
// "<data_root>/Data/SPDP/1stOrder_Rep_DP_G_<L>_T_<t>_dt_..._p1_<p1>_p2_<p2>_..._R_<r>.csv"
void terminal_percolationDornic_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_meas, double Rh0[], double t_max,
                                   double a, double b, double c, double D, double sigma, double dt, double dx, int r, int g);
void first_order_terminal_diff(int div, double t_event, double p_start, double p_end, double a, double b, double c, double D, double sigma,
                               double dt, double dx, int r, int g);
// frame capture at t_event (P/:2678, :2997; driver frameDP.cpp): the lattices of r replicates of the cubic DP model appended to
// "<data_root>/Data/SPDP/Frames/1RBOKCHOYFRAME_P_c_G_<L>_T_<t_event>_dt_..._R_<r>.csv"; capture_frame_decay sweeps a from one random frame
void frames_expanded_percolationDornic_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_meas, double Rh0[], double t_max,
                                          double t_event, double a, double b, double c, double D, double sigma, double dt, double dx, int g, int r);
void capture_frame_decay(double Rho_0[], int div, double t_max, double t_event, double a_start, double a_end, double b, double c, double D,
                         double sigma, double dt, double dx, int r, int g);
// finite-size scaling at a_c (P/:1424, :1538): percolationDornic_2D from a full frame for `div` grid sizes -- linspace(g_st, g_end)
// with the log10 time bins, resp. lnspace(g_st, g_end) with 30 linear bins in the reference's per-size stationary window (L > 30) --
// grid sizes dealt over the visible GPUs; CSV "<data_root>/Data/SPDP/DEBUG_P_c_Finite_Scaling_{Collapse,Stationary}_T_..._R_<r>.csv"
// with the first seven columns of the runner's rows, in grid order.
std::vector<double> lnspace(double start_in, double end_in, int log_points);   // P/:281
void critical_exp_finite_scaling_collapse(int div, int g_st, int g_end, double t_max, double a_c, double b, double D, double sigma,
                                          double dt, double dx, int r);
void critical_exp_finite_scaling_stationary(int div, int g_st, int g_end, double t_lim, double t_max, double a_c, double b, double D,
                                            double sigma, double dt, double dx, int r);
void first_order_critical_exp_delta(double Rho_0[], int div, double t_max, double a_start, double a_end, double b,
                                    double c, double D, double sigma, double dt, double dx, int r, int g);

// spreading experiments (theta, z exponents): every replicate starts from Rh0 (a single seed at (c_x, c_y)), stops when
// an active site reaches the border (theborderwall P/:380) or the lattice is empty; rows {thread, a, t, <N>_r, <R^2>_r, #surv}
void init_solitarytear(double array[], int length);                                                     // P/:90
void spreading_incremental_surv_runs(double t_N_R2_rep[][4], double X_curr[][3], int size);             // P/:230
void spreading_exp_Dornic_2D(std::vector<std::vector<double>> &Rho, std::vector<double> &t_meas, double Rh0[], double t_max,
                             double a, double b, double D, double sigma, double dt, double dx, int r, int g, int c_x, int c_y);
void critical_exp_theta_delta(double Rho_0[], int div, double t_max, double a_c, double b, double D, double sigma, double dt,
                              double dx, int r, int g);                                                  // P/:1667

// ---- Rietkerk, SPB = 1 (Sp = 3 fields: P, W, O) ---------------------------------------------
constexpr int Sp = 3, SpB = 1;
// frame_tmeas (R/ header :130): snapshot times; frames are written by save_frame() in the
// reference's CSV format.  Empty (default) = no snapshots.
extern std::vector<double> frame_tmeas;
extern std::string frame_folder; // default "<data_root>/Data/Rietkerk/Frames/Stochastic/1Sp/"
// Initial frame of replicate j (3*g*g doubles, fields P,W,O).  Default: the MFT speckle rule of
// init_exprtk_randbiMFTframe (R/:396-505) with the order_1stoc_test_rietkerk.cpp:144-160 expressions.
extern std::function<void(double *frame, int g, int replicate, double a, double a_c, double dP, const double clow[])> init_frame;

// Noise-off runner, R/:1816-2166: classical RK4 over the full right-hand side f_2D (R/:1718), r replicates, measured at
// every step while t <= 1 and then in the t_meas windows; rows {a, r, L, t, <W>, <O>, <P>, var, #surviving, #active}.
// Initial frame of replicate j: init_frame_noise_off (3*g*g doubles P, W, O); default = the reference's speckle rule
// init_randconstframe (R/:272: every cell is c_high with probability 0.015, else c_low, all three fields together) around
// the analytic steady state (R/:1855-1866), the mask drawn from options().seed instead of mt19937_64(random_device).
extern std::function<void(double *frame, int g, int replicate, double perc, const double c_high[], const double c_low[])> init_frame_noise_off;
// The runner's snapshots (R/:1959-1972): every replicate's frame at t = 0, for 500 <= t <= 2500 and in the last 15 measurement rows,
// "<data_root>/Data/Rietkerk/Frames/Non-Stochastic/<a_st>-<a_end>_dP_<dP>/FRAME_RAND_Basic_P_c_DP_G_<L>_T_<t>_dt_..._R_<j>.csv"; as
// in the reference the folder is the driver main's to create (no folder: an error line per frame, nothing written).
extern bool write_noise_off_frames;
void RK4_Wrapper_2D(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c, double gmax, double alpha,
                    double d, double rW, double W0, double (&D)[Sp], double (&K)[3], double a_st, double a_end, double dt, double dx,
                    double dP, int r, int g);
// noise-off a-sweep, R/:2170-2345 (driver order_test_rietkerk.cpp:182): RK4_Wrapper_2D per a-value, one CSV
// "<data_root>/Data/Rietkerk/Non-Stochastic/1stOrderCC_Rietkerk_P_c_Delta_DP_G_<L>_T_<T>_dt_..._R_<r>.csv" at setprecision(16)
void first_order_critical_exp_delta(int div, double t_max, double a_start, double a_end, double c, double gmax, double alpha, double d,
                                    double rW, double W0, double (&D)[Sp], double (&K)[3], double dt, double dx, double dP, int r, int g);
void rietkerk_Dornic_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);
void save_frame(D2Vec_Double &Rho_t, const std::string &parendir, const std::string &filenamePattern, double a,
                double a_st, double a_end, double t, double dt, double dx, double dP, int r, int g,
                std::string header = "", bool overwrite = false);                 // R/:2836
void save_prelimframe(D2Vec_Double &Rho_t, const std::string &parendir, const std::string &filenamePattern, double a,
                      double a_st, double a_end, double t, double dt, double dx, double dP, int r, int g,
                      std::string header = "", bool overwrite = false, bool delete_previous = false); // R/:2720
// the a-sweep: t_measure = {0} U log10 bins, frame_tmeas with the linear windows, all a-values x replicates as one
// batch of trials, final table to <stat_prefix><L>_T_..._R_<r>.csv at setprecision(16) (R/:4379-4617)
void first_order_critical_exp_delta_stochastic_MultiSp(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                       double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                       double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                       double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                       int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                       double Gstar = -1.0, double Vstar = -1.0);
// init_exprtk_randbiMFTframe (R/:396-505) for any number of fields, given the values expr[2 * n_fields] of the driver's MFT expressions
// at this a (below-a_c state first, coexistence state second); perc = 0.015 in the runners (R/:3403).  What the SPB >= 2 drivers assign
// to spb2::init_frame / spb3::init_frame.
void init_randbiMFTframe(double *frame, int n_fields, int g, int replicate, double a, double a_c, double dP, double perc,
                         const double clow[], const double expr[]);
// init_gaussframe (F/:1089-1199): the deterministic initial frame Rietkerk_FPE's runner uses as shipped (F/:3954-3962): Gaussian bumps
// of amplitude amp[s] for the first n_gauss (= SpB) fields, constants amp[s] for the rest; the runner passes amp = clow[Sp..2Sp)
void init_gaussframe(double *frame, int n_fields, int n_gauss, int g, const double amp[]);
// the sweeps' schedules (R/:4395-4459): returns t_measure and sets frame_tmeas
std::vector<double> sweep_time_schedule(double t_max, double dt);
// Rietkerk_FPE's runner and sweep of the same vegetation-only model (Rietkerk_FPE/rietkerk_bjork_FKE.cpp:3697, :4665 compiled with
// -DSPB=1): same signatures; clamp `v < 1e-10 && v >= 0` (F/:4460), checkpoints at the 85 / 90 / 95 / 100 % bins only (F/:3032), the FPE
// frame-saving rule (F/:2886, 2993), exit on a negative density (F/:4359).  BASELINE config 3's program at SPB = 1.
void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, std::vector<double> &t_meas, double t_max, double a, double c,
                                double gmax, double alpha, double rW, double W0, double (&D)[Sp], double (&v)[SpB],
                                double (&K)[3], double sigma[], double a_st, double a_end, double a_c,
                                double (&A)[SpB][SpB], double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB],
                                double pR[], int (&dtV)[SpB], double clow[], double dt, double dx, double dP, int r,
                                int g, double Gstar = -1.0, double Vstar = -1.0);
void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&v)[SpB], double (&K)[3], double sigma[], double (&A)[SpB][SpB],
                                                           double (&H)[SpB][SpB], double (&E)[SpB], double (&M)[SpB], double pR[],
                                                           int (&dtV)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar = -1.0, double Vstar = -1.0);
// burn-in frames (the reference's -DINIT=2 path): read_frame_csv parses one frame in save_frame()'s layout
// (init_csvconstframe R/:909, header matching of init_exprtk_readCSVcolumns_frame R/:609); use_burnin_frames() makes
// every replicate start from "<parendir><filenamePattern><r>.csv" (init_burnin_wrapper R/:836).
int read_frame_csv(const std::string &filename, int g, double *frame);
// ... for n_fields = 3 | 4 | 5 fields (P[, G[, Pr]], W, O), the columns located by header name (R/:665-705); and the burn-in rule for
// any of them as an init_frame function (what use_burnin_frames and its spb2 / spb3 twins assign)
int read_frame_csv_fields(const std::string &filename, int g, int n_fields, double *frame);
std::function<void(double *, int, int, double, double, double, const double[])> burnin_frame_rule(const std::string &parendir,
                                                                                                   const std::string &filenamePattern, int n_fields);
void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern);
void use_default_init_frame();   // back to the MFT speckle rule
// set_Prefix (R/:113-135): prefix and the three output locations derived from it, "<data_root>/Data/Rietkerk/..." for the reference's "../Data/Rietkerk/..."
void set_Prefix(const std::string &user_prefix, double mG = 0.0, double mP = 0.0);
// the same rule for a binary compiled with -DSPB=n_biota, writing the three locations into the caller's variables (spb2 / spb3 set_Prefix)
void set_Prefix_for(int n_biota, const std::string &user_prefix, double mG, double mP, std::string &frames, std::string &prelims,
                    std::string &stats);
void recursive_dir_create(const std::string &path);    // R/: mkdir -p
extern std::string prefix, prelim_folder, stat_prefix; // R/ header :97-108; empty folders = "<data_root>/Data/Rietkerk/..."
extern bool write_prelims;                             // PRELIM_AGGRAND_*.csv per replicate (R/:4327-4349) and the TimeSeries/PRELIM_TSERIES_* checkpoints (R/:3029)
extern const std::string prelimheader;                 // rietkerk_bjork_constants_1Sp.h:8-9

} // namespace tm_b200
