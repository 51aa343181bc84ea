/*
 * oracle/ref_harness_dp.cpp -- TEST INFRASTRUCTURE (see tm_oracle.h).
 *
 * C-callable doors into the UNMODIFIED reference DP integrator
 * (/root/reference/simulations/Percolation_FD/stochasticSPDP.cpp, compiled where it lies by
 * oracle/Makefile into oracle/_ref/).  Nothing here re-implements reference logic: it converts
 * plain arrays to the reference's argument types, arms the replay URBG, and calls the replicate
 * runners percolationDornic_2D (stochasticSPDP.cpp:940) / expanded_percolationDornic_2D (:1809).
 */
#include "SPDP.h" /* reference header, -I/root/reference/simulations/Percolation_FD */
#include <fcntl.h>

/* 12-argument overload defined at stochasticSPDP.cpp:940 (SPDP.h:96 still shows the old one) */
void percolationDornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[],
    double t_max, double a, double b, double D, double sigma, double dt, double dx, int r, int g);

namespace {
struct quiet { /* the runners chatter on stdout/stderr; park both on /dev/null for the call */
    int o, e, n;
    quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
              n = open("/dev/null", O_WRONLY); o = dup(1); e = dup(2); dup2(n, 1); dup2(n, 2); }
    ~quiet() { fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
               dup2(o, 1); dup2(e, 2); close(o); close(e); close(n); }
};
}

extern "C" {

/* Runs r replicates of one a-value.  rows_out[k] = {t, <rho>, var, n_surv, n_active, last} per
 * logged row (the runner's own output rows, stochasticSPDP.cpp:1130-1142 / 2139-2149).
 * u != NULL arms replay: the reference's std:: samplers consume u[0..n_u) in order.
 * Returns the number of rows, or -1. */
int ref_dp_run(int cubic, int L, double a, double b, double c, double D, double sigma, double dt,
               double dx, int r, const double *rh0, const double *t_meas_in, int n_tmeas,
               const double *u, uint64_t n_u, double *rows_out, int rows_cap,
               uint64_t *consumed, uint64_t *overflow)
{
    vector<vector<double>> Rho;
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    vector<double> Rh0(rh0, rh0 + (size_t)L * L);
    double t_max = t_meas.back();
    tm_replay_arm(u, n_u);
    {
        quiet q;
        if (cubic) expanded_percolationDornic_2D(Rho, t_meas, Rh0.data(), t_max, a, b, c, D, sigma, dt, dx, r, L);
        else       percolationDornic_2D(Rho, t_meas, Rh0.data(), t_max, a, b, D, sigma, dt, dx, r, L);
    }
    if (consumed) *consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    int n = (int)Rho.size();
    if (n > rows_cap) return -1;
    const int off = cubic ? 3 : 2; /* {a,(c),L,t,mean,var,surv,active,last} */
    for (int k = 0; k < n; k++)
        for (int j = 0; j < 6; j++) rows_out[6 * k + j] = Rho[k][off + j];
    return n;
}

/* spreading_exp_Dornic_2D (stochasticSPDP.cpp:1145): r replicates from rh0, seed site (c_x, c_y).
 * rows_out[k] = {thread, a, t, <N>_r, <R2>_r, n_surv} -- the runner's own output rows (:1294-1298). */
int ref_dp_spread_run(int L, double a, double b, double D, double sigma, double dt, double dx, int r,
                      const double *rh0, const double *t_meas_in, int n_tmeas, int c_x, int c_y,
                      const double *u, uint64_t n_u, double *rows_out, int rows_cap,
                      uint64_t *consumed, uint64_t *overflow)
{
    vector<vector<double>> Rho;
    vector<double> t_meas(t_meas_in, t_meas_in + n_tmeas);
    vector<double> Rh0(rh0, rh0 + (size_t)L * L);
    double t_max = t_meas.back();
    tm_replay_arm(u, n_u);
    {
        quiet q;
        spreading_exp_Dornic_2D(Rho, t_meas, Rh0.data(), t_max, a, b, D, sigma, dt, dx, r, L, c_x, c_y);
    }
    if (consumed) *consumed = tm_replay.pos;
    if (overflow) *overflow = tm_replay.overflow;
    tm_replay_arm(nullptr, 0);
    int n = (int)Rho.size();
    if (n > rows_cap) return -1;
    for (int k = 0; k < n; k++)
        for (int j = 0; j < 6; j++) rows_out[6 * k + j] = Rho[k][j];
    return n;
}

/* logarithm10_time_bins, stochasticSPDP.cpp:336 */
int ref_dp_time_bins(double t_max, double dt, double *out, int cap)
{
    vector<double> v;
    { quiet q; v = logarithm10_time_bins(t_max, dt); }
    if ((int)v.size() > cap) return -1;
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

/* RK4_Integrate (pointwise cubic remainder), stochasticSPDP.cpp:1775 */
void ref_dp_rk4(int L, double a, double b, double c, double D, double t, double dt, double dx, double *rho)
{
    vector<double> R(rho, rho + (size_t)L * L);
    { quiet q; RK4_Integrate(R, a, b, c, D, t, dt, dx, L); }
    for (size_t i = 0; i < R.size(); i++) rho[i] = R[i];
}

/* percNoDiff_2D (stochasticSPDP.cpp:549): the noise-off runner (one replicate whatever r is).  rows_out[k] = the runner's own
 * rows {a, replicate, t, cell, rho} (:643).  Returns the number of rows, or -1.  Its arrays live on the stack (g*g*10 ints +
 * 4*g*g doubles): keep g small. */
int ref_dp_noise_off(int L, double a, double b, double D, double dt, double dx, const double *rh0, const double *t_stop_in,
                     int n_tstop, double t_max, double *rows_out, int rows_cap)
{
    vector<vector<double>> Rho;
    vector<double> t_stop(t_stop_in, t_stop_in + n_tstop);
    vector<double> Rh0(rh0, rh0 + (size_t)L * L);
    { quiet q; percNoDiff_2D(Rho, t_stop, Rh0.data(), t_max, a, b, D, dt, dx, 1, L); }
    const int n = (int)Rho.size();
    if (n > rows_cap) return -1;
    for (int k = 0; k < n; k++)
        for (int j = 0; j < 5; j++) rows_out[5 * k + j] = Rho[k][j];
    return n;
}

}
