// tm_host_capi.cpp -- flat C doors into the C++ host layer (tm_host.h) so the Python tests can
// drive the reference-signature runners through ctypes.  No logic of its own.
#include "tm_host.h"
#include "tm_host_fke.h"
#include "tm_host_2sp.h"
#include "tm_host_side.h"
#include "tm_host_3sp.h"
#include <cmath>
#include <cstdio>
#include <cstring>

using namespace tm_b200;

extern "C" {

void tmh_set_options(int device, unsigned long long seed, int strict, const char *data_root)
{
    options().device = device; options().n_devices = 1; options().device_ids.clear(); options().seed = seed; options().strict = strict != 0;
    if (data_root) options().data_root = data_root;
}

// the GPUs a sweep is spread over: `first` .. `first + count - 1` (count = 0: every visible GPU from `first` on), and the
// cap on trials per device context (0: from the free device memory)
void tmh_set_devices(int first, int count, int max_trials_per_ctx)
{
    options().device = first; options().n_devices = count; options().max_trials_per_ctx = max_trials_per_ctx;
    options().device_ids.clear();
}

// explicit worker list (an id may repeat: two workers on one GPU); n = 0 clears it
void tmh_set_device_list(const int *ids, int n)
{
    options().device_ids.assign(ids, ids + (n > 0 ? n : 0));
}

unsigned long long tmh_stream_seed(unsigned long long seed, double a) { return stream_seed(seed, a); }

int tmh_time_bins(double t_max, double dt, double *out, int cap)
{
    std::vector<double> v = logarithm10_time_bins(t_max, dt);
    if ((int)v.size() > cap) return -1;
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

// rows_out[k][9]: the runner's output row, zero padded ({a,(c),L,t,mean,var,surv,active,last})
int tmh_dp_run(int cubic, int L, double a, double b, double c, double D, double sigma, double dt, double dx, int r,
               const double *rh0, const double *t_meas, int n_tmeas, double *rows_out, int cap)
{
    std::vector<std::vector<double>> Rho;
    std::vector<double> tm(t_meas, t_meas + n_tmeas), r0(rh0, rh0 + (size_t)L * L);
    try {
        if (cubic) expanded_percolationDornic_2D(Rho, tm, r0.data(), tm.back(), a, b, c, D, sigma, dt, dx, r, L);
        else       percolationDornic_2D(Rho, tm, r0.data(), tm.back(), a, b, D, sigma, dt, dx, r, L);
    } catch (const std::exception &) { return -2; }
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 9; j++) rows_out[9 * k + j] = j < Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}

// spreading experiment: rows_out[k][6] = {thread, a, t, <N>_r, <R^2>_r, #surviving}
int tmh_dp_spread_run(int L, double a, double b, double D, double sigma, double dt, double dx, int r, const double *rh0,
                      const double *t_meas, int n_tmeas, int c_x, int c_y, double *rows_out, int cap)
{
    std::vector<std::vector<double>> Rho;
    std::vector<double> tm(t_meas, t_meas + n_tmeas), r0(rh0, rh0 + (size_t)L * L);
    try {
        spreading_exp_Dornic_2D(Rho, tm, r0.data(), tm.back(), a, b, D, sigma, dt, dx, r, L, c_x, c_y);
    } catch (const std::exception &) { return -2; }
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 6; j++) rows_out[6 * k + j] = Rho[k][j];
    return (int)Rho.size();
}

// the theta/z driver (spreading_exp_test.cpp): writes ../Data/SPDP/SPREADING_P_c_ThDel_DP_G_*.csv under data_root
int tmh_dp_spread_sweep(int L, int div, double t_max, double a_c, double b, double D, double sigma, double dt, double dx, int r)
{
    std::vector<double> rho0((size_t)L * L, 0.0);
    try { critical_exp_theta_delta(rho0.data(), div, t_max, a_c, b, D, sigma, dt, dx, r, L); }
    catch (const std::exception &) { return -2; }
    return 0;
}

int tmh_dp_sweep(int cubic, int L, int div, double t_max, double a_start, double a_end, double b, double c, double D,
                 double sigma, double dt, double dx, int r)
{
    std::vector<double> rho0((size_t)L * L, 1.0); // init_fullframe, P/1storderDP_delta.cpp:45
    try {
        if (cubic) first_order_critical_exp_delta(rho0.data(), div, t_max, a_start, a_end, b, c, D, sigma, dt, dx, r, L);
        else       critical_exp_delta(rho0.data(), div, t_max, a_start, a_end, b, D, sigma, dt, dx, r, L);
    } catch (const std::exception &) { return -2; }
    return 0;
}

// Rietkerk SPB=1 with the order_1stoc_test_rietkerk.cpp constants; init = NULL -> default MFT speckle frame
int tmh_riet_run(int L, double a, double dt, int r, const double *init, const double *t_meas, int n_tmeas,
                 const double *frame_t, int n_frame_t, const char *frame_dir, double *rows_out, int cap)
{
    double c = 10000, gmax = 0.05e-3 / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0, dx = 0.1;
    double D[Sp] = {0.00025 / 24.0, 0.00025 / 24.0, 0.025 / 24.0};
    double K[3] = {0, 5, 5000};
    double sigma[Sp] = {sqrt(D[0] / (dx * dx)), 0, 0};
    double v[SpB] = {0}, A[SpB][SpB] = {{0}}, H[SpB][SpB] = {{0}}, E[SpB] = {1.0}, M[SpB] = {d}, pR[SpB] = {0.0};
    int dtV[SpB] = {0};
    double clow[2 * Sp] = {0, 1, 1, 5, 1, 1};
    std::vector<double> tm(t_meas, t_meas + n_tmeas);
    frame_tmeas.assign(frame_t, frame_t + n_frame_t);
    frame_folder = frame_dir ? frame_dir : "";
    write_prelims = false;
    auto saved = init_frame;
    const size_t n = (size_t)L * L;
    if (init) init_frame = [init, n](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + 3 * n * j, 3 * n * sizeof(double)); };
    D2Vec_Double Rho;
    int rc = 0;
    try {
        rietkerk_Dornic_2D_MultiSp(Rho, tm, tm.back(), a, c, gmax, alpha, rW, W0, D, v, K, sigma, a, a, 1.0 / 24.0, A, H, E, M, pR,
                                   dtV, clow, dt, dx, 10000.0, r, L);
    } catch (const std::exception &) { rc = -2; }
    init_frame = saved;
    if (rc) return rc;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 10; j++) rows_out[10 * k + j] = j < Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}

// the SPB=1 a-sweep with the order_1stoc_test_rietkerk.cpp constants; writes the reference's CSV tree under data_root
int tmh_riet_sweep(int L, int div, double t_max, double a_start, double a_end, double dt, int r)
{
    double c = 10000, gmax = 0.05e-3 / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0, dx = 0.1;
    double D[Sp] = {0.00025 / 24.0, 0.00025 / 24.0, 0.025 / 24.0};
    double K[3] = {0, 5, 5000};
    double sigma[Sp] = {sqrt(D[0] / (dx * dx)), 0, 0};
    double v[SpB] = {0}, A[SpB][SpB] = {{0}}, H[SpB][SpB] = {{0}}, E[SpB] = {1.0}, M[SpB] = {d}, pR[SpB] = {0.0};
    int dtV[SpB] = {0};
    double clow[2 * Sp] = {0, 1, 1, 5, 1, 1};
    frame_folder = ""; prelim_folder = ""; stat_prefix = ""; write_prelims = true;
    try {
        first_order_critical_exp_delta_stochastic_MultiSp(div, t_max, a_start, a_end, 1.0 / 24.0, c, gmax, alpha, rW, W0, D, v, K, sigma,
                                                          A, H, E, M, pR, dtV, clow, dt, dx, 10000.0, r, L);
    } catch (const std::exception &) { return -2; }
    return 0;
}

// ---- FKE movement (tm_host_fke.h) ------------------------------------------------------------------------------
// SPB = 2 doors: D[4], v[2], dtAdv[2] as order_2stoc_unity_rietkerk.cpp:202 passes them.
// out[7] = {beta_vel, sigma_D, sigma_vD.first, mu_vel_prefactor, bound_D, bound_vD.first, bound_vD.second} of species 1
void tmh_fke_set_params_2sp(const double *D_, const double *v_, const double *dtAdv_, double Dt, double dx, int L, int CI, double *out)
{
    double D[4] = {D_[0], D_[1], D_[2], D_[3]}, v[2] = {v_[0], v_[1]}, dtAdv[2] = {dtAdv_[0], dtAdv_[1]};
    set_global_FKE_params(D, v, dtAdv, Dt, dx, L, CI);     // the reference's call, array sizes deduced
    const FKEGlobals &G = fke_globals();
    out[0] = G.beta_vel[1]; out[1] = G.sigma_D[1]; out[2] = G.sigma_vD[1].first; out[3] = G.mu_vel_prefactor[1];
    out[4] = G.sig_D_ScaledBounds[1]; out[5] = G.sig_vD_ScaledBounds[1].first; out[6] = G.sig_vD_ScaledBounds[1].second;
}

void tmh_fke_override(int bound_D, int v_first, int v_second, double prefactor)
{
    FKEGlobals &G = fke_globals();
    G.sig_D_ScaledBounds[1] = bound_D; G.sig_vD_ScaledBounds[1] = {v_first, v_second}; G.mu_vel_prefactor[1] = prefactor;
}

void tmh_fke_precompute_gaussian(double sigma, int Rbound, double *out)
{
    const std::vector<double> s = precompute_gaussian(sigma, Rbound);
    for (int x = 0; x <= Rbound; x++) out[x] = s[x];
}

// advdiff_FKE_MultiSp with SPB = 2 containers (species 1 moves); returns 0 or -2 on an exception
int tmh_fke_advdiff_2sp(int g, const double *rho, const double *gamma, const double *v_eff, const double *gD, int nD,
                        const double *gV, int nV, double *out)
{
    const size_t n = (size_t)g * g;
    D2Vec_Double Rho_t(2, std::vector<double>(n, 0.0)), Rho_tsar(2, std::vector<double>(n, 0.0)), gam(2, std::vector<double>(n, 0.0));
    D2Vec_Pair_Double veff(2, std::vector<std::pair<double, double>>(n, {0.0, 0.0}));
    D2Vec_Double sD(2, std::vector<double>(1, 0.0)), sV(2, std::vector<double>(1, 0.0));
    for (size_t i = 0; i < n; i++) { Rho_t[1][i] = rho[i]; gam[1][i] = gamma[i]; veff[1][i] = {v_eff[2 * i], v_eff[2 * i + 1]}; }
    sD[1].assign(gD, gD + nD); sV[1].assign(gV, gV + nV);
    try { advdiff_FKE_MultiSp(Rho_t, Rho_tsar, gam, veff, sD, sV, g); } catch (const std::exception &) { return -2; }
    for (size_t i = 0; i < n; i++) out[i] = Rho_t[1][i];
    return 0;
}

// the -DBARRACUDA seam with caller-owned device pointers (one moving species)
int tmh_fke_seam_1sp(int L, int bound_D, int bvx, int bvy, double prefactor, void *d_rho, void *d_tsar, void *d_gamma, void *d_veff,
                     void *d_gD, void *d_gV)
{
    const int bD[1] = {bound_D}; const int2_ bV[1] = {{bvx, bvy}}; const double pf[1] = {prefactor}; const int off[1] = {0};
    if (copyToDeviceConstantMemory_AdvTerms(bD, bV, pf, off, off, 1) != 0) return -1;
    try {
        launchAdvDiffKernel_MultiSp((double *)d_rho, (double *)d_tsar, (double *)d_gamma, (std::pair<double, double> *)d_veff,
                                    (double *)d_gD, (double *)d_gV, L);
    } catch (const std::exception &) { return -2; }
    return 0;
}

// ---- burn-in frames -----------------------------------------------------------------------------------------------
// writes one 3-field frame with save_frame() (6-digit CSV) to <parendir><pattern><r>.csv
void tmh_save_frame(const char *parendir, const char *pattern, const double *frame, int g, int r, double a)
{
    const size_t n = (size_t)g * g;
    D2Vec_Double fr(Sp, std::vector<double>(n));
    for (int s = 0; s < Sp; s++) memcpy(fr[s].data(), frame + s * n, n * sizeof(double));
    save_frame(fr, parendir, pattern, a, a, a, 0.0, 0.1, 0.1, 0.0, r, g);
}
int tmh_read_frame_csv(const char *filename, int g, double *frame) { return read_frame_csv(filename, g, frame); }
int tmh_read_frame_csv_fields(const char *filename, int g, int n_fields, double *frame) { return read_frame_csv_fields(filename, g, n_fields, frame); }
// the burn-in rule of the SPB = 2 / 3 layers applied once: frame of `replicate` from "<parendir><pattern><r'>.csv"
int tmh_burnin_frame(int spb, const char *parendir, const char *pattern, int g, int replicate, double *frame)
{
    if (spb == 2) { tm_b200::spb2::use_burnin_frames(parendir, pattern); tm_b200::spb2::init_frame(frame, g, replicate, 0, 0, 0, nullptr); tm_b200::spb2::init_frame = nullptr; }
    else if (spb == 3) { tm_b200::spb3::use_burnin_frames(parendir, pattern); tm_b200::spb3::init_frame(frame, g, replicate, 0, 0, 0, nullptr); tm_b200::spb3::init_frame = nullptr; }
    else return -3;
    return 0;
}
// every later Rietkerk run starts from <parendir><pattern><r>.csv (parendir == NULL: back to the MFT speckle default)
void tmh_use_burnin(const char *parendir, const char *pattern)
{
    if (parendir && pattern) use_burnin_frames(parendir, pattern); else use_default_init_frame();
}

// the runner shims below return rows and frames to the caller: the runners' CSV side outputs are switched off for the call
// (no frame schedule, no prelim tables) unless the call is a sweep, which writes its whole tree under data_root
struct QuietSideOutputs {
    std::vector<double> ft; bool wp; bool on;
    explicit QuietSideOutputs(bool on_) : ft(frame_tmeas), wp(write_prelims), on(on_) { if (on) { frame_tmeas.clear(); write_prelims = false; } }
    ~QuietSideOutputs() { if (on) { frame_tmeas = ft; write_prelims = wp; } }
};

// SPB = 2 runner (tm_host_2sp.h).  par = {a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, sigma0, sigma1, v1, aij, hij, ej,
// m1, pR1, dt, dx} (the vector of tests/golden/make_runner2_golden.py); init: [n_init][4][L*L], replicate j starts from frame
// j % n_init; rows_out[k][14]; frames_out (may be NULL): [r][n_tmeas][4][L*L], the frame seen at every measurement bin.
int tmh_riet2_run(int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                  double *rows_out, int cap, double *frames_out)
{
    namespace s2 = tm_b200::spb2;
    double D[s2::Sp] = {par[9], par[10], par[11], par[12]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[s2::Sp] = {par[13], par[14], 0, 0};
    double v[s2::SpB] = {0, par[15]};
    double A[s2::SpB][s2::SpB] = {{0, par[16]}, {par[16], 0}}, H[s2::SpB][s2::SpB] = {{0, par[17]}, {par[17], 0}};
    double E[s2::SpB] = {1.0, par[18]}, M[s2::SpB] = {par[6], par[19]}, pR[s2::Sp] = {0.0, par[20]};
    int dtV[s2::SpB] = {0, 1};
    double clow[2 * s2::Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
    const size_t n4 = (size_t)4 * L * L;
    s2::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n4, n4 * sizeof(double)); };
    if (frames_out)
        s2::on_measure = [=](int j, int index, double, const double *f) { memcpy(frames_out + ((size_t)j * n_tmeas + index) * n4, f, n4 * sizeof(double)); };
    else s2::on_measure = nullptr;
    s2::D2Vec_Double Rho;
    std::vector<double> tm(t_meas, t_meas + n_tmeas);
    QuietSideOutputs quiet(true);
    try {
        s2::rietkerk_Dornic_2D_MultiSp(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                       1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[21], par[22], 10000.0, r, L);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_riet2_run: %s\n", e.what()); return -2; }
    s2::on_measure = nullptr;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 14; j++) rows_out[14 * k + j] = j < Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}

// the disc list of tm_b200::spb2::computeNeighboringSitesCentral as [n][2]; returns n (or -1 when cap is too small). Host only.
int tmh_central_sites(int R, int *out, int cap)
{
    const auto v = tm_b200::spb2::computeNeighboringSitesCentral(R);
    if ((int)v.size() > cap) return -1;
    for (size_t k = 0; k < v.size(); k++) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
    return (int)v.size();
}
// percNoDiff_2D: rows_out[k][5] = {a, replicate, t, cell, rho}
int tmh_dp_noise_off(int L, double a, double b, double D, double dt, double dx, const double *rh0, const double *t_stop, int n_tstop,
                     double t_max, double *rows_out, int cap)
{
    std::vector<std::vector<double>> Rho;
    std::vector<double> ts(t_stop, t_stop + n_tstop), r0(rh0, rh0 + (size_t)L * L);
    try { percNoDiff_2D(Rho, ts, r0.data(), t_max, a, b, D, dt, dx, 1, L); } catch (const std::exception &) { return -2; }
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 5; j++) rows_out[5 * k + j] = Rho[k][j];
    return (int)Rho.size();
}
// RK4_Wrapper_2D: par = {a, c, gmax, alpha, d, rW, W0, D0, D1, D2, K1, K2, dt, dx, dP}; init (may be NULL: the default speckle
// rule) = r frames [r][3][L*L]; rows_out[k][10]
int tmh_riet_det_run(const double *par, int L, int r, const double *init, const double *t_meas, int n_tmeas, double *rows_out, int cap)
{
    double D[Sp] = {par[7], par[8], par[9]};
    double K[3] = {0.0, par[10], par[11]};
    const size_t n3 = (size_t)3 * L * L;
    auto saved = init_frame_noise_off;
    if (init) init_frame_noise_off = [=](double *f, int, int j, double, const double[], const double[]) { memcpy(f, init + (size_t)j * n3, n3 * sizeof(double)); };
    D2Vec_Double Rho;
    std::vector<double> tm(t_meas, t_meas + n_tmeas);
    int rc = 0;
    const bool saved_frames = write_noise_off_frames;
    write_noise_off_frames = false;        // rows back to the caller: the runner's snapshots are switched off for the call
    try { RK4_Wrapper_2D(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], par[6], D, K, par[0], par[0], par[12], par[13], par[14], r, L); }
    catch (const std::exception &e) { fprintf(stderr, "tmh_riet_det_run: %s\n", e.what()); rc = -2; }
    write_noise_off_frames = saved_frames;
    init_frame_noise_off = saved;
    if (rc) return rc;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 10; j++) rows_out[10 * k + j] = Rho[k][j];
    return (int)Rho.size();
}
// the noise-off sweep (first_order_critical_exp_delta, Rietkerk): SPB = 1 driver constants, writes its CSV under data_root
int tmh_riet_det_sweep(int L, int div, double t_max, double a_start, double a_end, double dt, int r)
{
    double D[Sp] = {0.00025 / 24.0, 0.00025 / 24.0, 0.025 / 24.0}, K[3] = {0.0, 5.0, 5000.0};
    try {
        first_order_critical_exp_delta(div, t_max, a_start, a_end, 1e4, 0.05e-3 / 24, 0.2 / 24, 0.25 / 24, 0.2 / 24, 0.2, D, K, dt, 0.1, 10000.0, r, L);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_riet_det_sweep: %s\n", e.what()); return -2; }
    return 0;
}
// SPB = 2 a-sweep (tm_b200::spb2::first_order_critical_exp_delta_stochastic_MultiSp); par as tmh_riet2_run; writes its CSV
int tmh_riet2_sweep(int L, const double *par, int div, double t_max, double a_start, double a_end, int r, const double *init, int n_init)
{
    namespace s2 = tm_b200::spb2;
    double D[s2::Sp] = {par[9], par[10], par[11], par[12]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[s2::Sp] = {par[13], par[14], 0, 0};
    double v[s2::SpB] = {0, par[15]};
    double A[s2::SpB][s2::SpB] = {{0, par[16]}, {par[16], 0}}, H[s2::SpB][s2::SpB] = {{0, par[17]}, {par[17], 0}};
    double E[s2::SpB] = {1.0, par[18]}, M[s2::SpB] = {par[6], par[19]}, pR[s2::Sp] = {0.0, par[20]};
    int dtV[s2::SpB] = {0, 1};
    double clow[2 * s2::Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
    const size_t n4 = (size_t)4 * L * L;
    s2::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n4, n4 * sizeof(double)); };
    s2::on_measure = nullptr;
    try {
        s2::first_order_critical_exp_delta_stochastic_MultiSp(div, t_max, a_start, a_end, 1.0 / 24.0, par[1], par[2], par[3], par[4], par[5], D, v, K,
                                                              sigma, A, H, E, M, pR, dtV, clow, par[21], par[22], 10000.0, r, L);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_riet2_sweep: %s\n", e.what()); return -2; }
    return 0;
}

// Rietkerk_FPE, SPB = 2: rietkerk_DorFPE_2D_MultiSp (sweep != 0: its a-sweep over [a_start, a_end] in `div` values, CSV under data_root).
// par = the tmh_riet2_run vector plus par[23] = dtAdv[1]; set_global_FKE_params is called as the reference's driver does
// (order_2stoc_unity_rietkerk.cpp:202, CI = 4, Dt_analytic = dt).  rows_out / frames_out as tmh_riet2_run (ignored by the sweep).
int tmh_fpe2_run(int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                 double *rows_out, int cap, double *frames_out, int sweep, int div, double a_start, double a_end)
{
    namespace s2 = tm_b200::spb2;
    double D[s2::Sp] = {par[9], par[10], par[11], par[12]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[s2::Sp] = {par[13], par[14], 0, 0};
    double v[s2::SpB] = {0, par[15]};
    double A[s2::SpB][s2::SpB] = {{0, par[16]}, {par[16], 0}}, H[s2::SpB][s2::SpB] = {{0, par[17]}, {par[17], 0}};
    double E[s2::SpB] = {1.0, par[18]}, M[s2::SpB] = {par[6], par[19]}, pR[s2::Sp] = {0.0, par[20]};
    double dtAdv[s2::SpB] = {0.0, par[23]};
    int dtV[s2::SpB] = {0, 1};
    double clow[2 * s2::Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
    const size_t n4 = (size_t)4 * L * L;
    s2::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n4, n4 * sizeof(double)); };
    if (frames_out && !sweep)
        s2::on_measure = [=](int j, int index, double, const double *f) { memcpy(frames_out + ((size_t)j * n_tmeas + index) * n4, f, n4 * sizeof(double)); };
    else s2::on_measure = nullptr;
    s2::D2Vec_Double Rho;
    QuietSideOutputs quiet(!sweep);
    try {
        set_global_FKE_params(D, v, dtAdv, par[21], par[22], L, 4);
        if (sweep) {
            s2::first_order_critical_exp_delta_stochastic_MultiSp_FPE(div, t_meas[n_tmeas - 1], a_start, a_end, 1.0 / 24.0, par[1], par[2], par[3], par[4],
                                                                      par[5], D, v, K, sigma, A, H, E, M, pR, dtV, clow, par[21], par[22], 10000.0, r, L);
            return 0;
        }
        std::vector<double> tm(t_meas, t_meas + n_tmeas);
        s2::rietkerk_DorFPE_2D_MultiSp(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                       1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[21], par[22], 10000.0, r, L);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_fpe2_run: %s\n", e.what()); s2::on_measure = nullptr; return -2; }
    s2::on_measure = nullptr;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 14; j++) rows_out[14 * k + j] = j < Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}

// SPB = 3: tm_b200::spb3::rietkerk_Dornic_2D_MultiSp (sweep != 0: its a-sweep, CSV under data_root).  par = the 31-entry vector of
// tests/golden/make_runner3_golden.py {a, c, gmax, alpha, rW, W0, d, K1, K2, D0..D4, sigma0..2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2,
// pR1, pR2, dt, dx}; init: n_init frames [5][L*L], replicate j starts from frame j % n_init; rows_out[k][18]; frames_out (may be NULL):
// [r][n_tmeas][5][L*L].
int tmh_riet3_run_any(int fpe, int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                      double *rows_out, int cap, double *frames_out, int sweep, int div, double a_start, double a_end);
int tmh_riet3_run(int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                  double *rows_out, int cap, double *frames_out, int sweep, int div, double a_start, double a_end)
{
    return tmh_riet3_run_any(0, L, par, r, init, n_init, t_meas, n_tmeas, rows_out, cap, frames_out, sweep, div, a_start, a_end);
}
// Rietkerk_FPE, SPB = 3: par[31], par[32] = dtAdv[1], dtAdv[2]; set_global_FKE_params is called as the driver does (CI = 4, Dt = dt)
int tmh_fpe3_run(int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                 double *rows_out, int cap, double *frames_out, int sweep, int div, double a_start, double a_end)
{
    return tmh_riet3_run_any(1, L, par, r, init, n_init, t_meas, n_tmeas, rows_out, cap, frames_out, sweep, div, a_start, a_end);
}
int tmh_riet3_run_any(int fpe, int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                      double *rows_out, int cap, double *frames_out, int sweep, int div, double a_start, double a_end)
{
    namespace s3 = tm_b200::spb3;
    double D[s3::Sp] = {par[9], par[10], par[11], par[12], par[13]};
    double K[3] = {0.0, par[7], par[8]};
    double sigma[s3::Sp] = {par[14], par[15], par[16], 0, 0};
    double v[s3::SpB] = {0, par[17], par[18]};
    double A[s3::SpB][s3::SpB] = {{0, par[19], 0}, {par[19], 0, par[22]}, {0, par[22], 0}};
    double H[s3::SpB][s3::SpB] = {{0, par[20], 0}, {par[20], 0, par[23]}, {0, par[23], 0}};
    double E[s3::SpB] = {1.0, par[21], par[24]}, M[s3::SpB] = {par[6], par[25], par[26]}, pR[s3::Sp] = {0.0, par[27], par[28], 0.0, 0.0};
    int dtV[s3::SpB] = {0, 1, 1};
    double clow[2 * s3::Sp] = {0, 1, 1, 1, 1, 5, 1, 1, 1, 1};
    const size_t n5 = (size_t)5 * L * L;
    s3::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n5, n5 * sizeof(double)); };
    if (frames_out && !sweep)
        s3::on_measure = [=](int j, int index, double, const double *f) { memcpy(frames_out + ((size_t)j * n_tmeas + index) * n5, f, n5 * sizeof(double)); };
    else s3::on_measure = nullptr;
    s3::D2Vec_Double Rho;
    QuietSideOutputs quiet(!sweep);
    try {
        if (fpe) {
            double dtAdv[s3::SpB] = {0.0, par[31], par[32]};
            set_global_FKE_params(D, v, dtAdv, par[29], par[30], L, 4);
        }
        if (sweep) {
            if (fpe) s3::first_order_critical_exp_delta_stochastic_MultiSp_FPE(div, t_meas[n_tmeas - 1], a_start, a_end, 1.0 / 24.0, par[1], par[2], par[3], par[4],
                                                                               par[5], D, v, K, sigma, A, H, E, M, pR, dtV, clow, par[29], par[30], 10000.0, r, L);
            else s3::first_order_critical_exp_delta_stochastic_MultiSp(div, t_meas[n_tmeas - 1], a_start, a_end, 1.0 / 24.0, par[1], par[2], par[3], par[4],
                                                                       par[5], D, v, K, sigma, A, H, E, M, pR, dtV, clow, par[29], par[30], 10000.0, r, L);
            return 0;
        }
        std::vector<double> tm(t_meas, t_meas + n_tmeas);
        if (fpe) s3::rietkerk_DorFPE_2D_MultiSp(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                                1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[29], par[30], 10000.0, r, L);
        else s3::rietkerk_Dornic_2D_MultiSp(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                            1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[29], par[30], 10000.0, r, L);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_riet3_run: %s\n", e.what()); s3::on_measure = nullptr; return -2; }
    s3::on_measure = nullptr;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (size_t j = 0; j < 18; j++) rows_out[18 * k + j] = j < Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}
// frames_expanded_percolationDornic_2D (P/:2678) from the frame rh0[L*L]; the file lands under data_root
int tmh_dp_frames(int L, double a, double b, double c, double D, double sigma, double dt, double dx, int r, const double *rh0,
                  const double *t_meas, int n_tmeas, double t_event)
{
    std::vector<std::vector<double>> Rho;
    std::vector<double> tm(t_meas, t_meas + n_tmeas), r0(rh0, rh0 + (size_t)L * L);
    try { frames_expanded_percolationDornic_2D(Rho, tm, r0.data(), tm.empty() ? t_event : tm.back(), t_event, a, b, c, D, sigma, dt, dx, L, r); }
    catch (const std::exception &e) { fprintf(stderr, "tmh_dp_frames: %s\n", e.what()); return -2; }
    return 0;
}

// finite-size scaling drivers (P/:1424, :1538); kind 0 = collapse, 1 = stationary; CSV under data_root
int tmh_dp_finite_scaling(int kind, int div, int g_st, int g_end, double t_max, double a_c, double b, double D, double sigma, double dt,
                          double dx, int r)
{
    try {
        if (kind == 0) critical_exp_finite_scaling_collapse(div, g_st, g_end, t_max, a_c, b, D, sigma, dt, dx, r);
        else critical_exp_finite_scaling_stationary(div, g_st, g_end, 0.0, t_max, a_c, b, D, sigma, dt, dx, r);
    } catch (const std::exception &e) { fprintf(stderr, "tmh_dp_finite_scaling: %s\n", e.what()); return -2; }
    return 0;
}

// TEST SEAM (tm_host_2sp.h test_stepper): the SPB = 2 / SPB = 3 runners' bookkeeping and CSV writers over a Stepper made of three caller-
// supplied functions -- no device is touched.  tests/test_side_outputs.py passes the CPU oracle under sequential replay here and
// compares the output tree with the one the unmodified reference wrote.  spb = 2 | 3 (Rietkerk_FD runner), 12 | 13 (the Rietkerk_FPE
// runner rietkerk_DorFPE_2D_MultiSp of the same SPB); par / init / rows_out as tmh_riet2_run / tmh_riet3_run;
// frame_t: tm_b200::frame_tmeas for this call; folders "<root>/Frames/Stochastic/<n>Sp/<prefix>_" as set_Prefix lays them out.
typedef void (*tmh_upload_fn)(const double *frame);
typedef void (*tmh_run_fn)(uint32_t replicate, uint32_t step0, int nsteps);
typedef void (*tmh_download_fn)(double *frame, double *gamma, double *veff);
int tmh_side_run(int spb, int L, const double *par, int r, const double *init, int n_init, const double *t_meas, int n_tmeas,
                 const double *frame_t, int n_frame_t, const char *root, const char *user_prefix, double Gstar, double Vstar,
                 tmh_upload_fn up, tmh_run_fn run, tmh_download_fn down, double *rows_out, int cap, int prelims)
{
    struct Fn : tm_b200::side::Stepper {
        tmh_upload_fn u; tmh_run_fn r; tmh_download_fn d;
        void upload(const double *f) override { u(f); }
        void run(uint32_t j, uint32_t s0, int n) override { r(j, s0, n); }
        void download(double *f, double *g, double *v) override { d(f, g, v); }
    } fn;
    fn.u = up; fn.r = run; fn.d = down;
    const std::vector<double> saved_ft = frame_tmeas;
    const std::string saved_prefix = prefix;
    const bool saved_wp = write_prelims;
    frame_tmeas.assign(frame_t, frame_t + n_frame_t);
    prefix = user_prefix; write_prelims = prelims != 0;
    const bool fpe = spb >= 12;
    if (fpe) spb -= 10;
    const std::string sp = std::to_string(spb) + "Sp/";
    std::vector<double> tm(t_meas, t_meas + n_tmeas);
    std::vector<std::vector<double>> Rho;
    int rc = 0, width = 0;
    try {
        if (spb == 2) {
            namespace s2 = tm_b200::spb2;
            double D[s2::Sp] = {par[9], par[10], par[11], par[12]};
            double K[3] = {0.0, par[7], par[8]};
            double sigma[s2::Sp] = {par[13], par[14], 0, 0};
            double v[s2::SpB] = {0, par[15]};
            double A[s2::SpB][s2::SpB] = {{0, par[16]}, {par[16], 0}}, H[s2::SpB][s2::SpB] = {{0, par[17]}, {par[17], 0}};
            double E[s2::SpB] = {1.0, par[18]}, M[s2::SpB] = {par[6], par[19]}, pR[s2::Sp] = {0.0, par[20]};
            int dtV[s2::SpB] = {0, 1};
            double clow[2 * s2::Sp] = {0, 1, 1, 1, 5, 1, 1, 1};
            const size_t n4 = (size_t)4 * L * L;
            s2::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n4, n4 * sizeof(double)); };
            s2::on_measure = nullptr;
            s2::frame_folder = std::string(root) + "Frames/Stochastic/" + sp + user_prefix + "_";
            s2::prelim_folder = std::string(root) + "Prelims/Stochastic/" + sp + user_prefix + "_";
            s2::test_stepper = &fn;
            (fpe ? s2::rietkerk_DorFPE_2D_MultiSp : s2::rietkerk_Dornic_2D_MultiSp)(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                           1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[21], par[22], 10000.0, r, L, Gstar, Vstar);
            width = 14;
        } else if (spb == 3) {
            namespace s3 = tm_b200::spb3;
            double D[s3::Sp] = {par[9], par[10], par[11], par[12], par[13]};
            double K[3] = {0.0, par[7], par[8]};
            double sigma[s3::Sp] = {par[14], par[15], par[16], 0, 0};
            double v[s3::SpB] = {0, par[17], par[18]};
            double A[s3::SpB][s3::SpB] = {{0, par[19], 0}, {par[19], 0, par[22]}, {0, par[22], 0}};
            double H[s3::SpB][s3::SpB] = {{0, par[20], 0}, {par[20], 0, par[23]}, {0, par[23], 0}};
            double E[s3::SpB] = {1.0, par[21], par[24]}, M[s3::SpB] = {par[6], par[25], par[26]}, pR[s3::Sp] = {0.0, par[27], par[28], 0.0, 0.0};
            int dtV[s3::SpB] = {0, 1, 1};
            double clow[2 * s3::Sp] = {0, 1, 1, 1, 1, 5, 1, 1, 1, 1};
            const size_t n5 = (size_t)5 * L * L;
            s3::init_frame = [=](double *f, int, int j, double, double, double, const double[]) { memcpy(f, init + (size_t)(j % n_init) * n5, n5 * sizeof(double)); };
            s3::on_measure = nullptr;
            s3::frame_folder = std::string(root) + "Frames/Stochastic/" + sp + user_prefix + "_";
            s3::prelim_folder = std::string(root) + "Prelims/Stochastic/" + sp + user_prefix + "_";
            s3::test_stepper = &fn;
            (fpe ? s3::rietkerk_DorFPE_2D_MultiSp : s3::rietkerk_Dornic_2D_MultiSp)(Rho, tm, tm.back(), par[0], par[1], par[2], par[3], par[4], par[5], D, v, K, sigma, par[0], par[0],
                                           1.0 / 24.0, A, H, E, M, pR, dtV, clow, par[29], par[30], 10000.0, r, L, Gstar, Vstar);
            width = 18;
        } else rc = -3;
    } catch (const std::exception &e) { fprintf(stderr, "tmh_side_run: %s\n", e.what()); rc = -2; }
    tm_b200::spb2::test_stepper = nullptr; tm_b200::spb3::test_stepper = nullptr;
    tm_b200::spb2::frame_folder.clear(); tm_b200::spb2::prelim_folder.clear();
    tm_b200::spb3::frame_folder.clear(); tm_b200::spb3::prelim_folder.clear();
    frame_tmeas = saved_ft; prefix = saved_prefix; write_prelims = saved_wp;
    if (rc) return rc;
    if ((int)Rho.size() > cap) return -1;
    for (size_t k = 0; k < Rho.size(); k++)
        for (int j = 0; j < width; j++) rows_out[(size_t)width * k + j] = j < (int)Rho[k].size() ? Rho[k][j] : 0.0;
    return (int)Rho.size();
}
}
