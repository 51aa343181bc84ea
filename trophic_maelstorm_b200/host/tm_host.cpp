// tm_host.cpp -- see tm_host.h.  Bookkeeping follows the reference line by line (citations inline);
// all numerics happen behind the C ABI on the GPU.
#include "tm_host.h"
#include "../../include/tm_lattice.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <sys/stat.h>
#include <dirent.h>

namespace tm_b200 {

using std::vector;

Options &options()
{
    static Options o = [] {
        Options x;
        if (const char *e = getenv("TM_DEVICE")) { x.device = atoi(e); x.n_devices = 1; }   // one named GPU
        if (const char *e = getenv("TM_DEVICES")) x.n_devices = atoi(e);                    // GPUs device .. device + n - 1 (0: all visible)
        if (const char *e = getenv("TM_MAX_TRIALS_PER_CTX")) x.max_trials_per_ctx = atoi(e);
        if (const char *e = getenv("TM_DEVICE_LIST")) {                                        // explicit worker list, e.g. 0,1,2,3
            std::stringstream ss(e); std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) x.device_ids.push_back(atoi(tok.c_str()));
        }
        if (const char *e = getenv("TM_SEED")) x.seed = strtoull(e, nullptr, 0);
        if (const char *e = getenv("TM_STRICT")) x.strict = atoi(e) != 0;
        if (const char *e = getenv("TM_DATA_ROOT")) x.data_root = e;
        if (const char *e = getenv("TM_FILE_IDS")) x.file_ids = atoi(e) != 0;
        return x;
    }();
    return o;
}

unsigned long long stream_seed(unsigned long long seed, double a)
{
    unsigned long long bits;
    memcpy(&bits, &a, sizeof bits);
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (bits | 1ull);     // splitmix64 finaliser over (seed, a)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- the ensemble scheduler ---------------------------------------------------------------------------
// The reference spreads the a-values of a sweep over OpenMP threads (`#pragma omp parallel for`, R/:4509-4542,
// P/:2462-2488) and runs the replicates of an a-value one after the other.  Here the independent units -- a-values for the DP
// sweeps, contiguous chunks of (a-value, replicate) trials for the Rietkerk sweep -- are dealt round-robin to the visible GPUs:
// one host thread and one tm_ctx per GPU at a time, no device-to-device traffic, and all order-dependent host arithmetic
// (replicate-order running mean/variance, row order of the CSVs) done after the join, so the outputs do not depend on the
// number of GPUs.
thread_local int tl_device = -1;      // the GPU of the calling worker thread (-1: options().device)

int current_device() { return tl_device >= 0 ? tl_device : options().device; }

std::vector<int> device_list()
{
    const Options &o = options();
    if (!o.device_ids.empty()) return o.device_ids;
    const int visible = tm_device_count();
    int n = o.n_devices > 0 ? o.n_devices : visible - o.device;
    if (n < 1) n = 1;
    std::vector<int> d;
    for (int i = 0; i < n && (visible == 0 || o.device + i < visible); i++) d.push_back(o.device + i);
    if (d.empty()) d.push_back(o.device);
    return d;
}

void for_each_on_devices(int n_items, const std::function<void(int)> &fn)
{
    const std::vector<int> devs = device_list();
    const int nd = (int)std::min<size_t>(devs.size(), (size_t)std::max(n_items, 1));
    if (nd <= 1 || tl_device >= 0) {          // one GPU, or already inside a worker: run in place
        const int saved = tl_device;
        if (tl_device < 0) tl_device = devs[0];
        try { for (int i = 0; i < n_items; i++) fn(i); } catch (...) { tl_device = saved; throw; }
        tl_device = saved;
        return;
    }
    std::exception_ptr err; std::mutex m;
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d] {
            tl_device = devs[d];
            try { for (int i = d; i < n_items; i += nd) fn(i); }
            catch (...) { std::lock_guard<std::mutex> g(m); if (!err) err = std::current_exception(); }
        });
    for (auto &t : th) t.join();
    if (err) std::rethrow_exception(err);
}

namespace {
struct Ctx { // RAII over tm_ctx; errors become exceptions like the reference's exit() paths
    tm_ctx *c = nullptr;
    Ctx(const tm_params &p, int n_trials)
    {
        if (tm_create(&p, n_trials, current_device(), &c) != TM_OK)
            throw std::runtime_error(std::string("tm_create: ") + tm_last_error(nullptr));
        tm_set_strict(c, options().strict ? 1 : 0);
    }
    ~Ctx() { tm_destroy(c); }
    void ck(int rc, const char *what) { if (rc != TM_OK) throw std::runtime_error(std::string(what) + ": " + tm_last_error(c)); }
};

void mkdirs(const std::string &path)
{
    std::string cur;
    for (size_t i = 0; i < path.size(); i++) {
        cur += path[i];
        if (path[i] == '/' || i + 1 == path.size()) mkdir(cur.c_str(), 0777);
    }
}

// The reference decides what happens at each step from the accumulated double `t` (t += dt).
// Plan = for every loop iteration k: is it a measurement row, and which t_meas index it matched.
struct StepPlan { vector<int> measure; vector<int> window; vector<double> t; };

// DP schedule, P/:1003-1025: measure every step while t <= 1, then at the t_meas windows.
StepPlan plan_dp(const vector<double> &t_meas, double dt)
{
    StepPlan p;
    const double t_max = t_meas.back();
    double t = 0; size_t index = 0;
    while (t < t_max + dt) {
        const bool win = index < t_meas.size() && t >= t_meas[index] - dt / 2.0 && t < t_meas[index] + dt / 2.0;
        const bool meas = (t <= 1.0) || win;
        p.measure.push_back(meas); p.window.push_back(win ? (int)index : -1); p.t.push_back(t);
        if (meas && t > 1.0) index += 1;
        t += dt;
    }
    return p;
}
} // namespace

// ---- bookkeeping ---------------------------------------------------------------------------------

vector<double> linspace(double start, double end, int num) // P/:254-278
{
    vector<double> v;
    if (num == 0) return v;
    if (num == 1) { v.push_back(start); return v; }
    const double delta = (end - start) / (num - 1);
    for (int i = 0; i < num - 1; ++i) v.push_back(start + delta * i);
    v.push_back(end);
    return v;
}

vector<double> logarithm10_time_bins(double t_max, double dt) // P/:336-378
{
    vector<double> logspaced;
    if (t_max <= 1) return logspaced;
    const double tend_one = static_cast<double>(floor(log10(t_max) * 10.)) / 10.0;
    double power = 0.0;
    while (power <= tend_one) {
        const double t_raw = pow(10, power);
        double t_fin = static_cast<int>(t_raw / dt) * dt;
        if (t_fin <= 1.0) t_fin = t_fin + dt;
        if (logspaced.size() > 0 && t_fin != logspaced[logspaced.size() - 1]) logspaced.push_back(t_fin);
        else if (logspaced.size() == 0) logspaced.push_back(t_fin);
        power += 0.04;
    }
    const int n = (int)logspaced.size();
    for (int i = n - 1; i > 0; i--)
        if (logspaced[i] < logspaced[i - 1] + dt / 2.0) logspaced.erase(logspaced.begin() + i);
    return logspaced;
}

void var_mean_incremental_surv_runs(double A[][5], double X[][2], int size) // P/:173-208
{
    for (int i = 0; i < size; i++) {
        int r = A[i][3] + 1;
        if (X[i][0] == 0.0) break;
        const double mean_prev = A[i][1];
        if (A[i][3] == 0.0) {
            A[i][1] = X[i][1]; A[i][2] = 0.0; A[i][4] = X[i][0];
        } else {
            A[i][1] = (X[i][1] + (r - 1) * A[i][1]) / r;
            A[i][4] = (X[i][0] + (r - 1) * A[i][4]) / r;
            A[i][2] = (A[i][1] - mean_prev) * (A[i][1] - mean_prev)
                    + (1 / (r - 1)) * ((r - 2) * A[i][2] + (X[i][1] - A[i][1]) * (X[i][1] - A[i][1])); // integer 1/(r-1), as the reference
        }
        A[i][3] = r;
    }
}

// ---- DP ------------------------------------------------------------------------------------------

namespace {
void dp_runner(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double a, double b, double c, bool cubic,
               double D, double sigma, double dt, double dx, int r, int g)
{
    if (t_meas.empty() || r < 1) throw std::invalid_argument("percolationDornic_2D: empty t_meas or r < 1");
    tm_params p{};
    p.model = cubic ? TM_MODEL_DP_CUBIC : TM_MODEL_DP;
    p.L = g; p.dt = dt; p.dx = dx; p.a = a; p.b = b; p.c = c; p.D = D; p.sigma = sigma;
    Ctx ctx(p, r);
    const size_t n = (size_t)g * g;
    vector<double> host(n * r);
    for (int j = 0; j < r; j++) memcpy(&host[n * j], Rh0, n * sizeof(double)); // P/:988-992
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");
    ctx.ck(tm_seed(ctx.c, stream_seed(options().seed, a), 0), "tm_seed");

    const int tot_iter = static_cast<int>(1.0 / dt) + 1 + (int)t_meas.size(); // P/:949
    vector<double> acc_(5 * (size_t)tot_iter, 0.0);
    double (*acc)[5] = reinterpret_cast<double (*)[5]>(acc_.data());
    int ind = 0;
    for (double t = 0; t <= 1; t += dt) { if (ind < tot_iter) acc[ind][0] = t; ind += 1; } // P/:961
    for (size_t i = 0; i < t_meas.size() && ind + (int)i < tot_iter; i++) acc[ind + i][0] = t_meas[i];
    vector<double> surv_rep(t_meas.size() + 1, 0.0);
    vector<double> surv_rep3(t_meas.size() + 1, 0.0);   // ... as it stands after the third replicate (the PRELIMINARY table below)

    // per-replicate series Rho_M[s] = {N_active, <rho>}; a replicate stops logging once extinct (P/:1021)
    vector<vector<double>> RhoM(r, vector<double>(2 * (size_t)tot_iter, 0.0));
    vector<int> s_of(r, 0);
    vector<char> alive(r, 1);
    const StepPlan plan = plan_dp(t_meas, dt);
    vector<tm_stats> st(r);
    int pending = 0;
    for (size_t k = 0; k < plan.t.size(); k++) {
        if (plan.measure[k]) {
            if (pending) { ctx.ck(tm_step(ctx.c, pending), "tm_step"); pending = 0; }
            ctx.ck(tm_measure(ctx.c, st.data()), "tm_measure");
            bool any = false;
            for (int j = 0; j < r; j++) {
                if (!alive[j]) continue;
                if (st[j].nonfinite) { // P/:2022-2060: NaN -> exit(3)
                    std::cerr << "RK4 BLEW IT: non-finite density in replicate " << j << " at t=" << plan.t[k] << std::endl;
                    exit(3);
                }
                if (plan.window[k] >= 0 && st[j].mean[0] > 0) { surv_rep[plan.window[k]] += 1; if (j <= 2) surv_rep3[plan.window[k]] += 1; } // P/:1008-1013
                if (s_of[j] < tot_iter) { RhoM[j][2 * s_of[j]] = (double)st[j].active[0]; RhoM[j][2 * s_of[j] + 1] = st[j].mean[0]; }
                s_of[j] += 1;
                if (st[j].active[0] == 0) alive[j] = 0; else any = true;
            }
            if (!any) break;
        }
        pending += 1; // the Dornic + remainder update of this iteration
    }
    // replicate-order accumulation, P/:1074-1098
    for (int j = 0; j < r; j++) {
        double (*X)[2] = reinterpret_cast<double (*)[2]>(RhoM[j].data());
        if (j == 0) {
            for (int i = 0; i < tot_iter; i++) {
                acc[i][1] = X[i][1]; acc[i][2] = 0.0; acc[i][4] = X[i][0];
                if (X[i][1] > 0) acc[i][3] = 1; else if (X[i][1] == 0) acc[i][3] = 0;
            }
        } else {
            var_mean_incremental_surv_runs(acc, X, tot_iter);
        }
        if (j == 2 && write_prelims) {
            // P/:1100-1128, :2101-2128: the running table after three replicates, default stream precision, t_max = the last bin
            std::stringstream L, tm, d3, p1, rini, Sm;
            L << g; tm << t_meas.back(); d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a; rini << j; Sm << std::setprecision(3) << sigma;
            mkdirs(options().data_root + "/Data/SPDP");
            std::ofstream out(options().data_root + "/Data/SPDP/PRELIMINARY_P_c_Delta_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a1_" + p1.str()
                              + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv");
            out << (cubic ? " a , c, L, t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt \n"
                          : " a , L, t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt \n");
            for (int i = 0; i < tot_iter; i++) {
                out << a << ",";
                if (cubic) out << c << ",";
                out << g << "," << acc[i][0] << "," << acc[i][1] << "," << acc[i][2] << "," << acc[i][3] << "," << acc[i][4] << ",";
                if (i < ind) out << j << std::endl; else out << surv_rep3[i - ind] << std::endl;
            }
        }
    }
    for (int i = 0; i < tot_iter; i++) { // P/:1130-1142, :2139-2149
        const double last = (i < ind) ? static_cast<double>(r) : surv_rep[i - ind];
        if (cubic) Rho.push_back({a, c, static_cast<double>(g), acc[i][0], acc[i][1], acc[i][2], acc[i][3], acc[i][4], last});
        else       Rho.push_back({a, static_cast<double>(g), acc[i][0], acc[i][1], acc[i][2], acc[i][3], acc[i][4], last});
    }
}
} // namespace

void percolationDornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double /*t_max*/, double a,
                          double b, double D, double sigma, double dt, double dx, int r, int g)
{
    dp_runner(Rho, t_meas, Rh0, a, b, 0.0, false, D, sigma, dt, dx, r, g); // t_max := t_meas.back(), P/:958
}

void expanded_percolationDornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double /*t_max*/,
                                   double a, double b, double c, double D, double sigma, double dt, double dx, int r, int g)
{
    dp_runner(Rho, t_meas, Rh0, a, b, c, true, D, sigma, dt, dx, r, g);
}

namespace {
void write_dp_csv(const std::string &name, const char *header, const vector<vector<double>> &vec, bool cubic)
{
    mkdirs(options().data_root + "/Data/SPDP");
    std::ofstream out(options().data_root + "/Data/SPDP/" + name);
    out << header;
    for (size_t i = 0; i < vec.size(); i++) {
        const vector<double> &v = vec[i];
        if (cubic) // P/:2535-2538
            out << std::setprecision(8) << v[0] << "," << v[1] << "," << v[2] << "," << v[3] << "," << std::setprecision(16) << v[4]
                << "," << std::setprecision(16) << v[5] << "," << v[6] << "," << std::setprecision(16) << v[7] << "," << v[8] << std::endl;
        else       // P/:1408-1411
            out << std::setprecision(8) << v[0] << "," << v[1] << "," << v[2] << "," << std::setprecision(16) << v[3] << ","
                << std::setprecision(16) << v[4] << "," << v[5] << "," << std::setprecision(16) << v[6] << "," << v[7] << std::endl;
    }
}
} // namespace

void critical_exp_delta(double Rho_0[], int div, double t_max, double a_start, double a_end, double b, double D,
                        double sigma, double dt, double dx, int r, int g) // P/:1309-1421
{
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    vector<vector<double>> vec;
    vector<vector<vector<double>>> per_a(a_space.size());
    for_each_on_devices((int)a_space.size(), [&](int i) { // the reference's `omp for` over a-values; here one batch of r trials each, one GPU per a-value at a time
        percolationDornic_2D(per_a[i], t_measure, Rho_0, t_max, a_space[i], b, D, sigma, dt, dx, r, g);
    });
    for (const auto &rows : per_a) vec.insert(vec.end(), rows.begin(), rows.end());
    std::stringstream L, tm, d3, p1, p2, rini, Dm, Sm;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a_start; p2 << std::setprecision(4) << a_end;
    rini << r; Dm << std::setprecision(3) << D; Sm << std::setprecision(3) << sigma;
    write_dp_csv("DEBUG_P_c_Delta_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a1_" + p1.str() + "_a2_" + p2.str()
                 + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv", // P/:1395
                 " a , L,  t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt \n", vec, false);
}

void first_order_critical_exp_delta(double Rho_0[], int div, double t_max, double a_start, double a_end, double b,
                                    double c, double D, double sigma, double dt, double dx, int r, int g) // P/:2422-2547
{
    for (int i = 0; i < g * g; i++) Rho_0[i] = 0.25; // init_constframe(Rho_0, 0.25, g*g), P/:2427
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    vector<vector<double>> vec;
    vector<vector<vector<double>>> per_a(a_space.size());
    for_each_on_devices((int)a_space.size(), [&](int i) {
        expanded_percolationDornic_2D(per_a[i], t_measure, Rho_0, t_max, a_space[i], b, c, D, sigma, dt, dx, r, g);
    });
    for (const auto &rows : per_a) vec.insert(vec.end(), rows.begin(), rows.end());
    std::stringstream L, coco, tm, d3, p1, p2, rini, Sm, dix;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a_start; p2 << std::setprecision(4) << a_end;
    rini << r; Sm << std::setprecision(3) << sigma; coco << std::setprecision(4) << c; dix << std::setprecision(2) << dx;
    write_dp_csv("1stOrder_P_c_Delta_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_c_" + coco.str() + "_a1_" + p1.str()
                 + "_a2_" + p2.str() + "_dx_" + dix.str() + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv",
                 " a , c , L,  t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt \n", vec, true);
}

// ---- terminal state after a fixed time from a range of initial densities (P/:2158-2420, :2548-2676; driver bifurc_prob.cpp) -------

// terminal_percolationDornic_2D, P/:2158: the cubic DP step from the frame Rh0, r replicates; <rho> and the active sites at t = 0 and
// in the window of t_meas[0] (P/:2216-2224), one row per replicate {a, c, L, j, t_meas.back(), <rho(0)>, <rho(t)>, #active(t)}
// (P/:2413-2417).  The reference keeps stepping for another 100 time units without looking at the lattice again (t_max = t_meas.back()
// + 100, P/:2178): not done here.  Streams: keyed by (TM_SEED, a, Rh0[0]) so that the sweep's initial densities are independent.
void terminal_percolationDornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double /*t_max*/, double a,
                                   double b, double c, double D, double sigma, double dt, double dx, int r, int g)
{
    if (t_meas.empty() || r < 1) throw std::invalid_argument("terminal_percolationDornic_2D: empty t_meas or r < 1");
    tm_params p{};
    p.model = TM_MODEL_DP_CUBIC;
    p.L = g; p.dt = dt; p.dx = dx; p.a = a; p.b = b; p.c = c; p.D = D; p.sigma = sigma;
    Ctx ctx(p, r);
    const size_t n = (size_t)g * g;
    vector<double> host(n * r);
    for (int j = 0; j < r; j++) memcpy(&host[n * j], Rh0, n * sizeof(double));
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");
    ctx.ck(tm_seed(ctx.c, stream_seed(stream_seed(options().seed, a), Rh0[0]), 0), "tm_seed");
    vector<tm_stats> st0(r), st1(r);
    ctx.ck(tm_measure(ctx.c, st0.data()), "tm_measure");                          // t == 0
    int steps = 0;                                                                // iterations before the window of t_meas[0] opens
    for (double t = 0; t < t_meas.back() + 100 + dt; t += dt, steps++)
        if (t != 0.0 && t >= t_meas[0] - dt / 2.0 && t < t_meas[0] + dt / 2.0) break;
    ctx.ck(tm_step(ctx.c, steps), "tm_step");
    ctx.ck(tm_measure(ctx.c, st1.data()), "tm_measure");
    for (int j = 0; j < r; j++) {
        if (st0[j].nonfinite || st1[j].nonfinite) { std::cerr << "RK4 BLEW IT: non-finite density in replicate " << j << std::endl; exit(3); }   // P/:2370-2398
        const bool dead0 = st0[j].active[0] == 0;                                 // extinct at t = 0: the replicate stops there (P/:2222)
        Rho.push_back({a, c, static_cast<double>(g), static_cast<double>(j), t_meas.back(), st0[j].mean[0],
                       dead0 ? 0.0 : st1[j].mean[0], dead0 ? 0.0 : (double)st1[j].active[0]});
    }
}

// first_order_terminal_diff, P/:2548: the runner above from constant frames p in linspace(p_start, p_end, div), dealt over the GPUs;
// CSV "<data_root>/Data/SPDP/1stOrder_Rep_DP_G_<L>_T_<t>_dt_..._R_<r>.csv"
void first_order_terminal_diff(int div, double t_event, double p_start, double p_end, double a, double b, double c, double D, double sigma,
                               double dt, double dx, int r, int g)
{
    const vector<double> p_space = linspace(p_start, p_end, div);
    vector<vector<vector<double>>> per_p(p_space.size());
    for_each_on_devices((int)p_space.size(), [&](int i) {
        vector<double> Rho_0((size_t)g * g, p_space[i]);                          // init_constframe, P/:46
        vector<double> t_measure{t_event};
        terminal_percolationDornic_2D(per_p[i], t_measure, Rho_0.data(), t_event, a, b, c, D, sigma, dt, dx, r, g);
    });
    std::stringstream L, coco, tm, d3, p1, p2, rini, Sm, dix, aoili;
    L << g; tm << t_event; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << p_start; p2 << std::setprecision(4) << p_end;
    rini << r; Sm << std::setprecision(3) << sigma; coco << std::setprecision(4) << c; dix << std::setprecision(2) << dx; aoili << std::setprecision(4) << a;
    mkdirs(options().data_root + "/Data/SPDP");
    std::ofstream out(options().data_root + "/Data/SPDP/1stOrder_Rep_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + aoili.str()
                      + "_c_" + coco.str() + "_p1_" + p1.str() + "_p2_" + p2.str() + "_dx_" + dix.str() + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv");
    out << "  a , c , L, r, t ,  <<Rho(x; 0)>_x>, <<Rho(x; t)>_x>, # Active Sites \n";
    for (const auto &rows : per_p)
        for (const vector<double> &v : rows)
            out << std::setprecision(8) << v[0] << "," << v[1] << "," << v[2] << "," << v[3] << "," << v[4] << "," << std::setprecision(16) << v[5]
                << "," << std::setprecision(16) << v[6] << "," << std::setprecision(16) << v[7] << std::endl;   // P/:2664-2667
}

// ---- frame capture at a given time (P/:2678-2995, :2997-3129; driver frameDP.cpp) --------------------------------------------------

// frames_expanded_percolationDornic_2D, P/:2678: the cubic DP step from Rh0, r replicates; the lattice of every replicate is appended to
// "<data_root>/Data/SPDP/Frames/1RBOKCHOYFRAME_P_c_G_<L>_T_<t_event>_dt_<dt>_a1_<a>_Sig_<sigma>_dx_<dx>_R_<r>.csv" (header + L*L rows
// "a,c,t,x,rho" at default precision, replicate 0 truncating the file) at the step whose accumulated t falls in the window of t_event --
// and, since the dump hangs on an `else if` behind the measurement test (P/:2763, 2791), only if that step is past t = 1 and not in a
// t_meas window.  Nothing is returned in Rho (the reference's accumulation is commented out, P/:2957-2993); it keeps stepping to
// t_event + 50 without looking at the lattice again: not done here.
void frames_expanded_percolationDornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double /*t_max*/, double t_event,
                                          double a, double b, double c, double D, double sigma, double dt, double dx, int g, int r)
{
    (void)Rho;
    if (r < 1) throw std::invalid_argument("frames_expanded_percolationDornic_2D: r < 1");
    int steps = -1, k = 0;
    double t_dump = 0;
    size_t index = 0;
    for (double t = 0; t < t_event + 50 + dt; t += dt, k++) {                     // P/:2755-2796
        const bool win = index < t_meas.size() && t >= t_meas[index] - dt / 2.0 && t < t_meas[index] + dt / 2.0;
        if (t <= 1.0 || win) { if (t > 1.0) index += 1; continue; }
        if (t >= t_event - dt / 2.0 && t < t_event + dt / 2.0) { steps = k; t_dump = t; break; }
    }
    if (steps < 0) return;                                                        // the window is hidden by a measurement step: no file
    tm_params p{};
    p.model = TM_MODEL_DP_CUBIC;
    p.L = g; p.dt = dt; p.dx = dx; p.a = a; p.b = b; p.c = c; p.D = D; p.sigma = sigma;
    Ctx ctx(p, r);
    const size_t n = (size_t)g * g;
    vector<double> host(n * r);
    for (int j = 0; j < r; j++) memcpy(&host[n * j], Rh0, n * sizeof(double));
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");
    ctx.ck(tm_seed(ctx.c, stream_seed(options().seed, a), 0), "tm_seed");
    ctx.ck(tm_step(ctx.c, steps), "tm_step");
    ctx.ck(tm_download(ctx.c, host.data()), "tm_download");
    std::stringstream L, tm, d3, p1, dX, Sm, repli;
    L << g; tm << t_event; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a; repli << r;
    dX << std::setprecision(3) << dx; Sm << std::setprecision(3) << sigma;
    mkdirs(options().data_root + "/Data/SPDP/Frames");
    std::ofstream out(options().data_root + "/Data/SPDP/Frames/1RBOKCHOYFRAME_P_c_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a1_" + p1.str()
                      + "_Sig_" + Sm.str() + "_dx_" + dX.str() + "_R_" + repli.str() + ".csv");
    for (int j = 0; j < r; j++) {                                                 // P/:2806-2813, once per replicate
        out << " a, c, t , x,  Rho(x; t)\n";
        for (size_t i = 0; i < n; i++) out << a << "," << c << "," << t_dump << "," << i << "," << host[n * j + i] << std::endl;
    }
}

// capture_frame_decay, P/:2997: the runner above for a in linspace(a_start, a_end, div) from ONE random frame N(1, 0.25) clipped at 0
// (init_randframe P/:53: mt19937_64(random_device) there, splitmix64 from TM_SEED with Box-Muller here), the a-values dealt over the GPUs
void capture_frame_decay(double Rho_0[], int div, double t_max, double t_event, double a_start, double a_end, double b, double c, double D,
                         double sigma, double dt, double dx, int r, int g)
{
    unsigned long long s = options().seed ^ 0x5DEECE66Dull;
    auto u01 = [&]() { unsigned long long z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                       z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return ((z >> 11) + 0.5) * 0x1.0p-53; };
    for (int i = 0; i < g * g; i++) {
        const double v = 1.00 + 0.25 * std::sqrt(-2.0 * std::log(u01())) * std::cos(6.283185307179586 * u01());
        Rho_0[i] = v < 0 ? 0.0 : v;
    }
    const vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    for_each_on_devices((int)a_space.size(), [&](int i) {
        vector<vector<double>> none;
        vector<double> tm_i = t_measure;
        frames_expanded_percolationDornic_2D(none, tm_i, Rho_0, t_max, t_event, a_space[i], b, c, D, sigma, dt, dx, g, r);
    });
}

// ---- finite-size scaling drivers (P/:1424-1665): percolationDornic_2D at a_c over a range of grid sizes ------------

// lnspace, P/:281-301
vector<double> lnspace(double start_in, double end_in, int log_points)
{
    vector<double> v;
    const double logMin = std::log(start_in), logMax = std::log(end_in);
    const double delta = (logMax - logMin) / (log_points - 1);
    if (log_points == 0) return v;
    if (log_points == 1) { v.push_back(start_in); return v; }
    for (int i = 0; i < log_points - 1; ++i) v.push_back(std::pow(std::exp(1), logMin + delta * i));
    v.push_back(end_in);
    return v;
}

namespace {
// both drivers write the first seven columns {a, L, t, mean, var, #surviving, #active} of the runner's rows, P/:1522-1526, 1653-1657
void write_finite_scaling_csv(const std::string &name, const vector<vector<double>> &vec)
{
    mkdirs(options().data_root + "/Data/SPDP");
    std::ofstream out(options().data_root + "/Data/SPDP/" + name);
    out << " a , L, t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt \n";
    for (const vector<double> &v : vec)
        out << std::setprecision(8) << v[0] << "," << v[1] << "," << std::setprecision(16) << v[2] << "," << std::setprecision(16) << v[3]
            << "," << v[4] << "," << std::setprecision(16) << v[5] << "," << v[6] << std::endl;
}
std::string finite_scaling_name(const char *kind, double t_max, double dt, double a_c, int g_st, int g_end, double sigma, int r)
{
    std::stringstream AC, tm, d3, g1, g2, rini, Sm;
    tm << t_max; d3 << std::setprecision(3) << dt; g1 << std::setprecision(4) << g_st; g2 << std::setprecision(4) << g_end;
    rini << r; Sm << std::setprecision(3) << sigma; AC << a_c;
    return std::string("DEBUG_P_c_Finite_Scaling_") + kind + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + AC.str() + "_L1_" + g1.str()
         + "_L2_" + g2.str() + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv";
}
} // namespace

void critical_exp_finite_scaling_collapse(int div, int g_st, int g_end, double t_max, double a_c, double b, double D, double sigma,
                                          double dt, double dx, int r) // P/:1424-1536
{
    const vector<double> gv = linspace(g_st, g_end, div);
    vector<int> g_space(gv.size());
    for (size_t i = 0; i < gv.size(); i++) g_space[i] = static_cast<int>(gv[i]);
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    vector<vector<vector<double>>> per_g(g_space.size());
    for_each_on_devices((int)g_space.size(), [&](int i) {   // the reference's `omp for` over grid sizes: one GPU per grid size at a time
        const int g = g_space[i];
        vector<double> Rho_0((size_t)g * g, 1.0);           // init_fullframe, P/:39
        vector<double> tm_i = t_measure;
        percolationDornic_2D(per_g[i], tm_i, Rho_0.data(), t_max, a_c, b, D, sigma, dt, dx, r, g);
    });
    vector<vector<double>> vec;                              // in grid order (the reference: in the order its threads finish)
    for (const auto &rows : per_g) vec.insert(vec.end(), rows.begin(), rows.end());
    write_finite_scaling_csv(finite_scaling_name("Collapse", t_max, dt, a_c, g_st, g_end, sigma, r), vec);
}

void critical_exp_finite_scaling_stationary(int div, int g_st, int g_end, double /*t_lim*/, double t_max, double a_c, double b, double D,
                                            double sigma, double dt, double dx, int r) // P/:1538-1665
{
    const vector<double> gv = lnspace(g_st, g_end, div);
    vector<int> g_space(gv.size());
    for (size_t i = 0; i < gv.size(); i++) g_space[i] = static_cast<int>(gv[i]);
    for (int g : g_space)   // P/:1589-1598 assigns the window for g > 30 only and reads it uninitialised below that
        if (g <= 30) throw std::invalid_argument("critical_exp_finite_scaling_stationary: the reference defines no time window for L <= 30");
    vector<vector<vector<double>>> per_g(g_space.size());
    for_each_on_devices((int)g_space.size(), [&](int i) {
        const int g = g_space[i];
        double t_start, t_end;                               // P/:1589-1598
        if (g <= 50) { t_start = 300; t_end = 900; }
        else if (g <= 100) { t_start = 1000; t_end = 2500; }
        else if (g <= 150) { t_start = 1250; t_end = 8000; }
        else if (g <= 200) { t_start = 1400; t_end = 10000; }
        else { t_start = 2000; t_end = 15000; }
        if (const char *e = getenv("TM_FSS_WINDOW_SCALE")) { const double f = atof(e); if (f > 0) { t_start *= f; t_end *= f; } }   // tests: shorter windows
        vector<double> t_measure = linspace(t_start, t_end, 30);
        vector<double> Rho_0((size_t)g * g, 1.0);
        percolationDornic_2D(per_g[i], t_measure, Rho_0.data(), t_end + 10, a_c, b, D, sigma, dt, dx, r, g);
    });
    vector<vector<double>> vec;
    for (const auto &rows : per_g) vec.insert(vec.end(), rows.begin(), rows.end());
    write_finite_scaling_csv(finite_scaling_name("Stationary", t_max, dt, a_c, g_st, g_end, sigma, r), vec);
}

// ---- DP spreading experiments (P/:1145-1305, :1667-1772) -----------------------------------------

void init_solitarytear(double array[], int length) // P/:90-97
{
    for (int i = 0; i < length * length; i++) array[i] = 0;
    const int lhalf = static_cast<int>(length / 2.0);
    array[length * lhalf + lhalf] = 1;
}

void spreading_incremental_surv_runs(double A[][4], double X[][3], int size) // P/:230-252
{
    for (int i = 0; i < size; i++) {
        const int r = static_cast<int>(A[i][3] + 1); // number of surviving replicates
        if (X[i][0] == 0.0) break;                   // non-surviving run: nothing logged from here on
        A[i][1] = (X[i][0] + (r - 1) * A[i][1]) / r;
        A[i][2] = (X[i][1] + (r - 1) * A[i][2]) / r;
        A[i][3] = r;
    }
}

void spreading_exp_Dornic_2D(vector<vector<double>> &Rho, vector<double> &t_meas, double Rh0[], double /*t_max*/, double a,
                             double b, double D, double sigma, double dt, double dx, int r, int g, int c_x, int c_y)
{
    if (t_meas.empty() || r < 1) throw std::invalid_argument("spreading_exp_Dornic_2D: empty t_meas or r < 1");
    tm_params p{};
    p.model = TM_MODEL_DP;
    p.L = g; p.dt = dt; p.dx = dx; p.a = a; p.b = b; p.c = 0.0; p.D = D; p.sigma = sigma;
    Ctx ctx(p, r);
    const size_t n = (size_t)g * g;
    vector<double> host(n * r);
    for (int j = 0; j < r; j++) memcpy(&host[n * j], Rh0, n * sizeof(double)); // P/:1188-1192
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");
    ctx.ck(tm_seed(ctx.c, stream_seed(options().seed, a), 0), "tm_seed");

    const int tot_iter = static_cast<int>(1.0 / dt) + 1 + (int)t_meas.size(); // P/:1153
    vector<double> acc_(4 * (size_t)tot_iter, 0.0);
    double (*acc)[4] = reinterpret_cast<double (*)[4]>(acc_.data());
    int ind = 0;
    for (double t = 0; t <= 1; t += dt) { if (ind < tot_iter) acc[ind][0] = t; ind += 1; } // P/:1163-1166
    for (size_t i = 0; i < t_meas.size() && ind + (int)i < tot_iter; i++) acc[ind + i][0] = t_meas[i];

    // per-replicate series Rho_M[s] = {N(t), R^2(t), 1}; replicates run side by side as trials, each one stops at its own
    // border breach (P/:1201-1212) or extinction (P/:1216-1217) -- later states of a stopped trial are simply ignored
    vector<vector<double>> RhoM(r, vector<double>(3 * (size_t)tot_iter, 0.0));
    vector<int> s_of(r, 0);
    vector<char> alive(r, 1);
    const StepPlan plan = plan_dp(t_meas, dt);
    vector<tm_spread> sp(r);
    int breaches = 0;
    for (size_t k = 0; k < plan.t.size(); k++) {
        ctx.ck(tm_measure_spread(ctx.c, c_x, c_y, sp.data()), "tm_measure_spread");
        bool any = false;
        for (int j = 0; j < r; j++) {
            if (!alive[j]) continue;
            if (sp[j].border) { alive[j] = 0; breaches += 1; continue; }
            if (plan.measure[k]) {
                if (sp[j].n_active == 0) { alive[j] = 0; continue; }
                if (s_of[j] < tot_iter) {
                    double *row = &RhoM[j][3 * (size_t)s_of[j]];
                    row[0] = (double)sp[j].n_active; row[1] = sp[j].sq_dist / (double)sp[j].n_active; row[2] = 1.0;
                }
                s_of[j] += 1;
            }
            any = true;
        }
        if (!any) break;
        ctx.ck(tm_step(ctx.c, 1), "tm_step");
    }
    for (int j = 0; j < r; j++) { // replicate-order accumulation, P/:1264-1287
        double (*X)[3] = reinterpret_cast<double (*)[3]>(RhoM[j].data());
        if (j == 0) {
            for (int i = 0; i < tot_iter; i++) {
                acc[i][1] = X[i][0]; acc[i][2] = X[i][1];
                if (X[i][0] > 0) acc[i][3] = 1; else if (X[i][0] == 0) acc[i][3] = 0;
            }
        } else {
            spreading_incremental_surv_runs(acc, X, tot_iter);
        }
    }
    for (int i = 0; i < tot_iter; i++) // P/:1294-1298; column 0 is the reference's OpenMP thread number
        Rho.push_back({0.0, a, acc[i][0], acc[i][1], acc[i][2], acc[i][3]});
    if (!options().quiet) std::cout << "IN SUMMATION WE HAVE A BREACH COUNT OF:\t" << breaches << std::endl;
}

void critical_exp_theta_delta(double Rho_0[], int div, double t_max, double a_c, double b, double D, double sigma, double dt,
                              double dx, int r, int g) // P/:1667-1772
{
    init_solitarytear(Rho_0, g);
    const int c_x = static_cast<int>(g / 2), c_y = c_x;
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    vector<vector<double>> vec;
    for (int i = 0; i < div; i++) { // the reference's `omp for` over supra-replicates: div batches of r trials each
        vector<vector<double>> rows;
        const unsigned long long seed0 = options().seed;
        options().seed = seed0 + 0x9E3779B97F4A7C15ull * (unsigned long long)i; // independent batches (the reference reseeds per replicate)
        spreading_exp_Dornic_2D(rows, t_measure, Rho_0, t_max, a_c, b, D, sigma, dt, dx, r, g, c_x, c_y);
        options().seed = seed0;
        vec.insert(vec.end(), rows.begin(), rows.end());
    }
    std::stringstream L, tm, d3, p1, rini, Sm;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << std::setprecision(5) << a_c; rini << r; Sm << std::setprecision(3) << sigma;
    mkdirs(options().data_root + "/Data/SPDP");
    std::ofstream out(options().data_root + "/Data/SPDP/SPREADING_P_c_ThDel_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str()
                      + "_a1_" + p1.str() + "_Sig_" + Sm.str() + "_R_" + rini.str() + ".csv"); // P/:1748-1749
    out << " #, a , t ,  <N(t)>_r, <MSD(N(t))>_r, <P(t)>_r \n";
    for (size_t i = 0; i < vec.size(); i++) // P/:1758-1760
        out << vec[i][0] << "," << vec[i][1] << "," << std::setprecision(16) << vec[i][2] << "," << std::setprecision(16) << vec[i][3]
            << "," << vec[i][4] << "," << vec[i][5] << std::endl;
}

// ---- Rietkerk SPB = 1 ----------------------------------------------------------------------------

std::vector<double> frame_tmeas;
std::string frame_folder;

// splitmix64 step for the file-name tokens below
static unsigned long long token_mix(unsigned long long &s)
{
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void independent_run_defaults()
{
    Options &o = options();
    if (getenv("TM_SEED")) return;                                            // a reproducible run: the caller's seed, constant tokens
    std::random_device rd;
    o.seed = ((unsigned long long)rd() << 32) ^ (unsigned long long)rd();
    if (!getenv("TM_FILE_IDS")) o.file_ids = true;
    std::cout << "TM_SEED=" << o.seed << "   (streams and file-name tokens of this run; set TM_SEED to repeat it)" << std::endl;
}

std::string jid_token(double a)
{
    if (!options().file_ids) return "0";
    unsigned long long s = stream_seed(options().seed ^ 0x7D1Dull, a);
    const double u = (double)(token_mix(s) >> 11) * (1.0 / 9007199254740992.0);
    std::stringstream t; t << (int)std::round(u * 10000);                     // R/:3634
    return t.str();
}

std::string id_token(double a, int replicate)
{
    if (!options().file_ids) return "0";
    unsigned long long s = stream_seed(options().seed ^ (0x1D0ull + 0x9E3779B97F4A7C15ull * (unsigned long long)(replicate + 1)), a);
    const double u = (double)(token_mix(s) >> 11) * (1.0 / 9007199254740992.0);
    std::stringstream t; t << (u * 1000.0) / 1000.0;                          // R/:4343, default stream precision
    return t.str();
}

namespace {
// splitmix64: the speckle mask of the default initial frame (the reference uses mt19937_64(random_device), R/:420)
inline unsigned long long splitmix(unsigned long long &s)
{
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void default_init_frame(double *f, int g, int replicate, double a, double a_c, double dP, const double clow[])
{
    // MFT expressions of order_1stoc_test_rietkerk.cpp:144-160 (the driver's strings), SPB = 1 constants :28-37
    const double c = 10000, gmax = 0.05e-3 / 24.0, d = 0.25 / 24.0, rW = 0.2 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, K1 = 5;
    const double Wstar = d * K1 / (c * gmax - d);
    const double perc = 0.015; // R/:3403
    const size_t n = (size_t)g * g;
    unsigned long long s = options().seed ^ (0xA5A5A5A5ull * (unsigned long long)(replicate + 1));
    double e[6]; // {V_prev, W_prev, O_prev, V, W, O}(a)
    e[0] = 0; e[1] = (1 / rW) * a; e[2] = (1 / (alpha * W0)) * a;
    e[3] = (c / d) * a + (-rW * Wstar * c / d); e[4] = Wstar * a - 0; e[5] = 120.4744181 * a + 0.58243061;
    if (a < a_c) { // R/:450-473
        const double high = std::max(dP, clow[0] * e[0]);
        for (size_t i = 0; i < n; i++) f[i] = ((splitmix(s) >> 11) * 0x1.0p-53 < perc) ? high : clow[0] * e[0];
        for (int sp = 1; sp < Sp; sp++) for (size_t i = 0; i < n; i++) f[sp * n + i] = clow[sp] * e[sp];
    } else {       // R/:474-496
        const double high = std::max(dP, clow[Sp] * e[Sp]);
        for (size_t i = 0; i < n; i++) f[i] = ((splitmix(s) >> 11) * 0x1.0p-53 < perc) ? high : e[Sp];
        for (int sp = 1; sp < Sp; sp++) for (size_t i = 0; i < n; i++) f[sp * n + i] = clow[sp + Sp] * e[sp + Sp];
    }
}
} // namespace

// init_exprtk_randbiMFTframe, R/:396-505, for any number of fields: `expr` holds the values of the driver's 2 * n_fields MFT expressions
// (MFT_Vec_CoexExpr: the state below a_c, then the coexistence state) at this a.  The speckle mask comes from options().seed and the
// replicate (splitmix64) where the reference seeds mt19937_64 from random_device (R/:420-423).
void init_randbiMFTframe(double *f, int n_fields, int g, int replicate, double a, double a_c, double dP, double perc, const double clow[],
                         const double expr[])
{
    const size_t n = (size_t)g * g;
    unsigned long long s = options().seed ^ (0xA5A5A5A5ull * (unsigned long long)(replicate + 1));
    const int off = a < a_c ? 0 : n_fields;                                      // R/:450, :474
    const double base = a < a_c ? clow[0] * expr[0] : expr[n_fields];            // R/:463, :487
    const double high = std::max(dP, clow[off] * expr[off]);                     // R/:452, :476
    for (size_t i = 0; i < n; i++) f[i] = ((splitmix(s) >> 11) * 0x1.0p-53 <= perc) ? high : base;
    for (int sp = 1; sp < n_fields; sp++) {
        const double v = expr[off + sp] == 0 ? clow[off + sp] : clow[off + sp] * expr[off + sp];   // R/:467-470, :491-494
        for (size_t i = 0; i < n; i++) f[(size_t)sp * n + i] = v;
    }
}

// init_gaussframe + init_gaussiantear, F/:1089-1199: what Rietkerk_FPE's runner starts from as shipped (F/:3954-3962, amp = clow[Sp..2Sp),
// sd = {g/8, g/16, g/16, ...}).  The first n_gauss fields are Gaussian bumps amp[s] exp(-((i - cx)^2 + (j - cy)^2) / (2 sd^2)) around the
// corners (g/8, g/8), (7g/8, g/8), (7g/8, 7g/8), (g/8, 7g/8) in turn (not wrapped periodically), the others the constants amp[s].
void init_gaussframe(double *f, int n_fields, int n_gauss, int g, const double amp[])
{
    const size_t n = (size_t)g * g;
    for (int s = 0; s < n_gauss; s++) {
        const double cx = (s % 4) % 3 == 0 ? int(g / 8) : int(7 * g / 8), cy = (s % 4) < 2 ? int(g / 8) : int(7 * g / 8);   // F/:1120-1121
        const double sd = s == 0 ? g / 8.0 : g / 16.0;                                                                     // F/:3959
        for (int i = 0; i < g; i++)
            for (int j = 0; j < g; j++)
                f[s * n + (size_t)i * g + j] = amp[s] * std::exp(-((i - cx) * (i - cx) + (j - cy) * (j - cy)) / (2 * sd * sd));   // F/:1196
    }
    for (int s = n_gauss; s < n_fields; s++)
        for (size_t i = 0; i < n; i++) f[s * n + i] = amp[s];
}

std::function<void(double *, int, int, double, double, double, const double[])> init_frame = default_init_frame;

void save_frame(D2Vec_Double &Rho_t, const std::string &parendir, const std::string &filenamePattern, double a, double,
                double, double, double, double, double, int r, int g, std::string header, bool) // R/:2836-2883
{
    std::stringstream rini; rini << r;
    mkdirs(parendir);
    std::ofstream frame_dp(parendir + filenamePattern + rini.str() + ".csv");
    if (header == "") header = "a_c,  x,  P(x; t), W(x; t), O(x; t) \n"; // rietkerk_bjork_constants_1Sp.h:6
    frame_dp << header;
    for (int i = 0; i < g * g; i++) {
        frame_dp << a << "," << i;
        for (size_t s = 0; s < Rho_t.size(); s++) frame_dp << "," << Rho_t[s][i]; // default (6-digit) precision, R/:2866
        frame_dp << std::endl;
    }
}

void save_prelimframe(D2Vec_Double &Rho_t, const std::string &parendir, const std::string &filenamePattern, double a, double,
                      double, double, double, double, double, int r, int g, std::string header, bool overwrite, bool) // R/:2720-2783
{
    mkdirs(parendir);
    std::string filename = parendir + filenamePattern + std::to_string(r) + ".csv";
    if (!overwrite) { // never overwrite: bump the replicate suffix (R/:2730-2740)
        int rep = r;
        struct stat sb;
        while (stat(filename.c_str(), &sb) == 0) filename = parendir + filenamePattern + std::to_string(++rep) + ".csv";
    }
    std::ostringstream oss;
    oss << (header != "" ? header : std::string("a_c,  x,  P(x; t), W(x; t), O(x; t) \n"));
    for (size_t i = 0; i < Rho_t.size(); i++) {
        oss << a << "," << r + 1 << "," << g << ",";
        for (size_t k = 0; k < Rho_t[i].size(); k++) oss << Rho_t[i][k] << ","; // default precision, trailing comma
        oss << "\n";
    }
    if (FILE *fp = fopen(filename.c_str(), "w")) { fprintf(fp, "%s", oss.str().c_str()); fclose(fp); }
    else std::cerr << "Error: Could not open PRELIM file " << filename << std::endl;
}

std::string prefix = "DiC-NREF-LI";                    // R/ header :97
std::string prelim_folder;                            // default "<data_root>/Data/Rietkerk/Prelims/Stochastic/1Sp/<prefix>_"
std::string stat_prefix;                              // default "<data_root>/Data/Rietkerk/Stochastic/1Sp/1stCC_Rietkerk_<prefix>_P_c_G_"
bool write_prelims = true;

void set_Prefix_for(int n_biota, const std::string &user_prefix, double mG, double mP, std::string &frames, std::string &prelims,
                    std::string &stats) // R/:113-135 (F/:143-165) for a binary compiled with -DSPB=n_biota
{
    prefix = user_prefix;
    mG = (mG == int(mG)) ? int(mG) : std::round(mG * 100) / 100;
    mP = (mP == int(mP)) ? int(mP) : std::round(mP * 100) / 100;
    std::ostringstream g_, p_;
    g_ << std::defaultfloat << mG; p_ << std::defaultfloat << mP;
    std::string sp = std::to_string(n_biota) + "Sp";
    if (mG > 0.0 && mP > 0.0) sp += "_mG_" + g_.str() + "_mP_" + p_.str();
    else if (mG > 0.0 && mP == 0.0) sp += "_mG_" + g_.str();
    const std::string root = options().data_root + "/Data/Rietkerk/";
    frames = root + "Frames/Stochastic/" + sp + "/" + prefix + "_";
    prelims = root + "Prelims/Stochastic/" + sp + "/" + prefix + "_";
    stats = root + "Stochastic/1stOrderCC_Rietkerk_" + prefix + "_P_c_G_";
}

void set_Prefix(const std::string &user_prefix, double mG, double mP) { set_Prefix_for(SpB, user_prefix, mG, mP, frame_folder, prelim_folder, stat_prefix); }

void recursive_dir_create(const std::string &path) { mkdirs(path); }
const std::string prelimheader = " a , r, L, t , <<W(x; t)>_x>_r, <<O(x; t)>_x>_r,  <<P(x; t)>_x>_r, Var[<P(x; t)>_x]_r, # Surviving Runs P(x; t),"
                                 " # Active Sites P(x; t), \n"; // rietkerk_bjork_constants_1Sp.h:8-9

namespace {
// All a-values of a sweep x r replicates as ONE batch of trials (the reference runs a-values on OpenMP threads
// and replicates serially, R/:4519, :3501); rows come back per a-value in the reference's layout.
void riet_batch(vector<D2Vec_Double> &Rho_per_a, const vector<double> &a_values, vector<double> &t_meas, double c, double gmax,
                double alpha, double rW, double W0, double (&D)[Sp], double (&K)[3], double sigma[], double a_st, double a_end,
                double a_c, double (&M)[SpB], double clow[], double dt, double dx, double dP, int r, int g, double Gstar, double Vstar,
                bool fpe = false)   // fpe: Rietkerk_FPE's runner of the same model (rietkerk_DorFPE_2D_MultiSp, F/:3697, -DSPB=1)
{
    if (t_meas.empty() || r < 1 || a_values.empty()) throw std::invalid_argument("rietkerk_Dornic_2D_MultiSp: empty t_meas, no a-value or r < 1");
    const int na = (int)a_values.size(), nt = na * r;
    tm_params p{};
    p.model = TM_MODEL_RIETKERK; p.L = g; p.dt = dt; p.dx = dx; p.a = a_values[0];
    p.rc = c; p.gmax = gmax; p.alpha = alpha; p.rW = rW; p.W0 = W0; p.d = M[0]; p.K1 = K[1]; p.K2 = K[2];
    p.D0 = D[0]; p.D1 = D[1]; p.D2 = D[2]; p.sigma0 = sigma[0];
    p.clamp_eps = 1.0e-12; p.clamp_lo = -1.0e-80; // R/:3402, :4176
    if (fpe) { p.clamp_eps = 1.0e-10; p.clamp_lo = 0.0; }   // F/:4460 (epsilon of the header :119)
    const size_t n = (size_t)g * g;

    const int tot_iter = (int)t_meas.size();               // R/:3438
    const int W_ = 4 * SpB + 3;                            // Sp4_1: {t, <W>, <O>, mean, var, surv, active}
    vector<vector<double>> RhoM(nt, vector<double>(2 * Sp * (size_t)tot_iter, 0.0)); // [index][2s] = N, [2s+1] = mean
    const double t_end = t_meas.back();
    const std::string fdir = frame_folder.empty() ? options().data_root + "/Data/Rietkerk/Frames/Stochastic/1Sp/" + prefix + "_" : frame_folder;
    const std::string pdir = prelim_folder.empty() ? options().data_root + "/Data/Rietkerk/Prelims/Stochastic/1Sp/" + prefix + "_" : prelim_folder;
    std::stringstream a1, a2, dimitri; a1 << a_st; a2 << a_end; dimitri << dP;
    std::string subdir = a1.str() + "-" + a2.str() + "_dP_" + dimitri.str();
    {   // R/:2907-2912, :4224, :4269, :4337: the equilibrium tag of the run folders
        std::stringstream geq, veq; geq << std::setprecision(5) << Gstar; veq << std::setprecision(5) << Vstar;
        if (Gstar != -1) subdir += "_Geq_" + geq.str(); else if (Vstar != -1) subdir += "_Veq_" + veq.str();
    }

    vector<char> broke(nt, 0);
    // save_prelimfileswrapper's table of ONE replicate, R/:3040-3079: var_mean_incremental_surv_runs(temp, Rho_M, index + 1, 0)
    // on a zeroed table = the vegetation rows up to its first empty bin, the water means everywhere; time in column 0
    auto write_tseries = [&](const vector<double> &X, const vector<double> &tm_, int idx, double t_now, double a, double a_st_, double a_end_,
                             double dt_, double dx_, double dP_, double D2, int j, int g_, const std::string &dir) {
        D2Vec_Double tab(idx + 1, vector<double>(W_, 0.0));
        for (int i = 0; i <= idx; i++) {
            const double N = X[2 * Sp * i + 0];
            if (N == 0.0) break;
            tab[i][3] = X[2 * Sp * i + 1]; tab[i][4] = 0.0; tab[i][6] = N; tab[i][5] = 1;
        }
        for (int i = 0; i <= idx; i++) {
            tab[i][1] = X[2 * Sp * i + 2 * (Sp - 2) + 1]; tab[i][2] = X[2 * Sp * i + 2 * (Sp - 1) + 1];
            tab[i][0] = tm_[i];
        }
        std::stringstream L, tm, d3, p1;
        L << g_; tm << t_now; d3 << std::setprecision(3) << dt_; p1 << std::setprecision(5) << a;
        (void)D2;
        const std::string pattern = "/PRELIM_TSERIES_P_c_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str()
                                  + "_jID_" + jid_token(a) + "_R_";   // R/ header :112, R/:3076 (jID = the thread's random ID there)
        save_prelimframe(tab, dir, pattern, a, a_st_, a_end_, t_now, dt_, dx_, dP_, j, g_, prelimheader, false, false);
    };

    // One chunk = the contiguous global trial ids [c0, c1) (trial id = a-index * r + replicate) in one device context on the
    // calling worker's GPU.  The Philox key is (seed, global trial id, step, cell) and the initial frame is drawn from the
    // global trial id, so a trial's trajectory does not depend on the chunking or on the GPU it lands on.
    auto run_chunk = [&](int c0, int c1) {
    const int nc = c1 - c0;
    Ctx ctx(p, nc);
    vector<double> host(3 * n * nc);
    for (int tr = c0; tr < c1; tr++) {
        ctx.ck(tm_set_trial_a(ctx.c, tr - c0, a_values[tr / r]), "tm_set_trial_a");
        init_frame(&host[3 * n * (tr - c0)], g, tr, a_values[tr / r], a_c, dP, clow);
    }
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");
    ctx.ck(tm_seed(ctx.c, options().seed, (uint32_t)c0), "tm_seed");
    vector<tm_stats> st(nc);
    double t = 0; int index = 0; size_t frame_index = 0; int pending = 0;
    while (t < t_end + dt) { // R/:3648
        const bool meas = index < tot_iter && (t == 0 || (t >= t_meas[index] - dt / 2.0 && t < t_meas[index] + dt / 2.0)); // R/:3659
        const bool snap = frame_index < frame_tmeas.size()
                       && (t == 0 || (t >= frame_tmeas[frame_index] - dt / 2.0 && t < frame_tmeas[frame_index] + dt / 2.0)); // R/:3736
        if ((meas || snap) && pending) { ctx.ck(tm_step(ctx.c, pending), "tm_step"); pending = 0; }
        if (meas) {
            ctx.ck(tm_measure(ctx.c, st.data()), "tm_measure");
            for (int j = c0; j < c1; j++) {
                if (broke[j]) continue;   // this replicate left the time loop at an earlier checkpoint: its later rows stay 0 (R/:3714)
                if (st[j - c0].nonfinite) { std::cerr << "RHO NAAN PARATHA in trial " << j << " at t=" << t << "\n"; exit(3); } // R/:4206-4276
                if (fpe && st[j - c0].negative) { std::cerr << "RHO FALLS BELOW O in trial " << j << " at t=" << t << "\n"; exit(3); }   // F/:4359-4369 (checked here at the measurement steps)
                for (int s = 0; s < Sp; s++) {
                    RhoM[j][2 * Sp * index + 2 * s] = (double)st[j - c0].active[s];
                    RhoM[j][2 * Sp * index + 2 * s + 1] = st[j - c0].mean[s];
                }
            }
            // per-replicate time-series checkpoints, R/:3710-3716 -> save_prelimfileswrapper R/:3029-3140
            if (((!fpe && index == int(tot_iter * 0.6)) || index == int(tot_iter * 0.85) || index == int(tot_iter * 0.9)
                                  || index == int(tot_iter * 0.95) || index == int(tot_iter - 1))) {   // F/:3032: no 60 % mark
                for (int tr = c0; tr < c1; tr++) {
                    if (broke[tr]) continue;
                    const double a = a_values[tr / r]; const int j = tr % r;
                    const bool dead = RhoM[tr][2 * Sp * index + 0] == 0.0;       // no vegetation left: R/:3087
                    if (write_prelims) write_tseries(RhoM[tr], t_meas, index, t, a, a_st, a_end, dt, dx, dP, D[2], j, g, pdir + subdir + "/TimeSeries");
                    if (dead) {   // the future checkpoints are written now, with zeros after this time, and the replicate stops
                        const int range[5] = {int(tot_iter * 0.6), int(tot_iter * 0.85), int(tot_iter * 0.9), int(tot_iter * 0.95), int(tot_iter - 1)};
                        for (int k = fpe ? 1 : 0; k < 5; k++) {
                            if (index > range[k] || !write_prelims) continue;
                            write_tseries(RhoM[tr], t_meas, range[k], t_meas[range[k]], a, a_st, a_end, dt, dx, dP, D[2], j, g,
                                          pdir + subdir + "/TimeSeries");
                        }
                        broke[tr] = 1;
                    }
                }
            }
            index += 1;
        }
        if (snap) {
            const size_t ftot = frame_tmeas.size();
            bool save = (frame_index + 10 >= ftot) || (t >= 50000 && t <= 150000) || (t >= (fpe ? 50 : 20) && t <= 15000) || t == 0; // R/:3739, F/:2886
            if (fpe && t > 1500 && t < 6000 && frame_index % 2 != 0) save = false;   // F/:2993-3004
            if (save) {
                ctx.ck(tm_download(ctx.c, host.data()), "tm_download");
                for (int tr = c0; tr < c1; tr++) {
                    if (broke[tr]) continue;   // the reference left this replicate's time loop (R/:3714): no further frames of it
                    const double a = a_values[tr / r]; const int j = tr % r;
                    D2Vec_Double fr(Sp, vector<double>(n));
                    for (int s = 0; s < Sp; s++) memcpy(fr[s].data(), &host[(3 * (size_t)(tr - c0) + s) * n], n * sizeof(double));
                    std::stringstream L, tm, d3, p1, Dm1, dix;
                    L << g; tm << t; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a; dix << std::setprecision(2) << dx;
                    Dm1 << std::setprecision(3) << D[1];
                    const std::string pattern = "/FRAME_P_c_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str()
                                              + "_D1_" + Dm1.str() + "_jID_" + jid_token(a) + "_dx_" + dix.str() + "_R_"; // R/:2914
                    save_frame(fr, fdir + subdir, pattern, a, a_st, a_end, t, dt, dx, dP, j, g);
                }
            }
            frame_index += 1;
        }
        pending += 1;
        t += dt;
    }
    };  // run_chunk

    // chunks: an equal share per GPU, capped by what fits in the free memory of a GPU (2 ping-pong states + the rho* scratch
    // field per trial) and by options().max_trials_per_ctx; dealt round-robin, a GPU works through its chunks in order
    {
        const int ndev = (int)device_list().size();
        long long cap = (nt + ndev - 1) / ndev;
        uint64_t need = 0, free_b = 0, total_b = 0;
        if (tm_ctx_bytes(&p, 1, &need) == TM_OK && tm_device_mem(current_device(), &free_b, &total_b) == TM_OK && need > 0) {
            const long long fit = (long long)((free_b * 0.9) / (double)need);
            if (fit >= 1 && fit < cap) cap = fit;
        }
        if (options().max_trials_per_ctx > 0 && options().max_trials_per_ctx < cap) cap = options().max_trials_per_ctx;
        if (cap < 1) cap = 1;
        const int n_chunks = (int)((nt + cap - 1) / cap);
        for_each_on_devices(n_chunks, [&](int k) {
            const long long c0 = (long long)nt * k / n_chunks, c1 = (long long)nt * (k + 1) / n_chunks;
            run_chunk((int)c0, (int)c1);
        });
    }
    const double t = t_end;   // the reference's loop variable after the time loop (file names below)
    // per a-value: replicate-order accumulation (R/:4294-4322, var_mean_incremental_surv_runs R/:1293-1343)
    Rho_per_a.assign(na, D2Vec_Double());
    for (int ia = 0; ia < na; ia++) {
        const double a = a_values[ia];
        vector<vector<double>> acc(tot_iter, vector<double>(W_, 0.0));
        for (int i = 0; i < tot_iter; i++) acc[i][0] = t_meas[i];
        for (int j = 0; j < r; j++) {
            const vector<double> &X = RhoM[ia * r + j];
            if (j == 0) {
                for (int i = 0; i < tot_iter; i++) {
                    const double N = X[2 * Sp * i + 0], mean = X[2 * Sp * i + 1];
                    acc[i][3] = mean; acc[i][4] = 0.0; acc[i][6] = N;
                    if (N > 0) acc[i][5] = 1; else if (N == 0) acc[i][5] = 0;
                    acc[i][1] = X[2 * Sp * i + 2 * (Sp - 2) + 1]; acc[i][2] = X[2 * Sp * i + 2 * (Sp - 1) + 1];
                }
            } else {
                for (int i = 0; i < tot_iter; i++) { // vegetation: surviving runs only, stops at the first dead bin (R/:1306-1310)
                    const double N = X[2 * Sp * i + 0], mean = X[2 * Sp * i + 1];
                    const int rr = (int)(acc[i][5] + 1);
                    if (N == 0.0) break;
                    const double mean_prev = acc[i][3];
                    if (acc[i][5] == 0.0) { acc[i][3] = mean; acc[i][4] = 0.0; acc[i][6] = N; }
                    else {
                        acc[i][3] = (mean + (rr - 1) * acc[i][3]) / rr;
                        acc[i][6] = (N + (rr - 1) * acc[i][6]) / rr;
                        acc[i][4] = (acc[i][3] - mean_prev) * (acc[i][3] - mean_prev)
                                  + (1 / (rr - 1)) * ((rr - 2) * acc[i][4] + (mean - acc[i][3]) * (mean - acc[i][3])); // int 1/(r-1), R/:1325
                    }
                    acc[i][5] = rr;
                }
                for (int i = 0; i < tot_iter; i++) { // soil and surface water: plain running mean over all replicates (R/:1333-1337)
                    acc[i][1] = (X[2 * Sp * i + 2 * (Sp - 2) + 1] + j * acc[i][1]) / (j + 1);
                    acc[i][2] = (X[2 * Sp * i + 2 * (Sp - 1) + 1] + j * acc[i][2]) / (j + 1);
                }
            }
            if (write_prelims) { // the running table after every replicate, R/:4327-4349 (jID is a random number there)
                std::stringstream L, tm, d3, p1, Dm;
                L << g; tm << t_end; d3 << std::setprecision(3) << dt; p1 << std::setprecision(5) << a; Dm << std::setprecision(4) << D[2];
                const std::string pattern = "/PRELIM_AGGRAND_P_c_ID_" + id_token(a, j) + "_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str()
                                          + "_D2_" + Dm.str() + "_R_";
                save_prelimframe(acc, pdir + subdir, pattern, a, a_st, a_end, t, dt, dx, dP, j, g, prelimheader, true);
            }
        }
        for (int i = 0; i < tot_iter; i++) { // R/:4361-4372: {a, r, L, t, <W>, <O>, <P>, var, surv, active}
            vector<double> row = {a, static_cast<double>(r), static_cast<double>(g)};
            row.insert(row.end(), acc[i].begin(), acc[i].end());
            Rho_per_a[ia].push_back(row);
        }
    }
}

// switchsort_and_bait, R/:1409-1436 (insert_type "linspace", erase = true as the sweep calls it)
vector<double> switchsort_and_bait(vector<double> vals, double low, double high, int num_points)
{
    vals.erase(std::remove_if(vals.begin(), vals.end(), [low, high](double x) { return x >= low && x <= high; }), vals.end());
    vector<double> lin = linspace(low, high, num_points);
    vals.insert(vals.end(), lin.begin(), lin.end());
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    return vals;
}
} // namespace

// ---- RK4_Wrapper_2D R/:1816-2166: the noise-off Rietkerk runner ------------------------------------------------------
static void default_init_frame_noise_off(double *f, int g, int replicate, double perc, const double c_high[], const double c_low[])
{
    const size_t n = (size_t)g * g;
    unsigned long long s = options().seed ^ (0xD37ull + 0x9E3779B97F4A7C15ull * (unsigned long long)(replicate + 1));
    for (size_t i = 0; i < n; i++) {
        const double u = (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0);
        for (int k = 0; k < Sp; k++) f[k * n + i] = (u <= perc) ? c_high[k] : c_low[k];             // R/:283-290
    }
}
std::function<void(double *, int, int, double, const double[], const double[])> init_frame_noise_off = default_init_frame_noise_off;

bool write_noise_off_frames = true;

// The snapshot the noise-off runner saves at a measurement row (R/:1959-1972, through save_frame R/:2836): the name grammar of R/:1962-1971
// with its stream precisions, the reference's header for this runner, rows `a,i,P,W,O` at the default 6-digit precision.  As in the
// reference the folder is NOT created here (the driver main makes "<a_st>-<a_end>_dP_<dP>" with the default precision of dP, the runner
// names it with setprecision(3): for dP = 10000 the two differ and the reference writes no frames, only its error line) and an existing
// file is never overwritten: the replicate suffix becomes (largest suffix present) + 1 (findMaxRepNo R/:1544).
static void save_noise_off_frame(const double *frame, double a, double a_st, double a_end, double t, double dt, double dx, double dP,
                                 double D0, double D1, double alpha, double gmax, int j, int g)
{
    std::stringstream L, tm, d3, p1, rini, gm, a1, a2, Dm0, Dm1, alph, dix, dimitri;
    L << g; tm << t; d3 << std::setprecision(3) << dt; p1 << std::setprecision(4) << a; dix << std::setprecision(2) << dx;
    rini << j; Dm0 << std::setprecision(3) << D0; Dm1 << std::setprecision(3) << D1; gm << std::setprecision(3) << gmax;
    alph << std::setprecision(3) << alpha; dimitri << std::setprecision(3) << dP; a1 << a_st; a2 << a_end;
    const std::string parendir = options().data_root + "/Data/Rietkerk/Frames/Non-Stochastic/" + a1.str() + "-" + a2.str() + "_dP_" + dimitri.str();
    const std::string pattern = "FRAME_RAND_Basic_P_c_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str() + "_D0_" + Dm0.str()
                              + "_D1_" + Dm1.str() + "_dx_" + dix.str() + "_alph_" + alph.str() + "_gmax_" + gm.str() + "_R_";
    std::string filename = parendir + "/" + pattern + rini.str() + ".csv";
    struct stat sb;
    if (stat(filename.c_str(), &sb) == 0) {
        int max_rep = 0;
        if (DIR *dir = opendir(parendir.c_str())) {
            while (struct dirent *e = readdir(dir)) {
                const std::string name = e->d_name;
                if (name.size() <= pattern.size() + 4 || name.compare(0, pattern.size(), pattern) != 0 || name.compare(name.size() - 4, 4, ".csv") != 0) continue;
                const std::string num = name.substr(pattern.size(), name.size() - 4 - pattern.size());
                if (num.find_first_not_of("0123456789") != std::string::npos) continue;
                const int rep = atoi(num.c_str());
                if (rep > max_rep) max_rep = rep;
            }
            closedir(dir);
        }
        filename = parendir + "/" + pattern + std::to_string(max_rep + 1) + ".csv";
    }
    const size_t n = (size_t)g * g;
    std::ostringstream oss;
    oss << "a_c,  x,  P(x; tmax), W(x; tmax), O(x; tmax) \n";
    for (size_t i = 0; i < n; i++) oss << a << "," << i << "," << frame[i] << "," << frame[n + i] << "," << frame[2 * n + i] << "\n";
    if (FILE *fp = fopen(filename.c_str(), "w")) { fprintf(fp, "%s", oss.str().c_str()); fclose(fp); }
    else std::cerr << "Error: Could not open file " << filename << std::endl;
}

void RK4_Wrapper_2D(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax, double alpha,
                    double d, double rW, double W0, double (&D)[Sp], double (&K)[3], double a_st, double a_end, double dt,
                    double dx, double dP, int r, int g)
{
    if (t_meas.empty() || r < 1) throw std::invalid_argument("RK4_Wrapper_2D: empty t_meas or r < 1");
    const int n_meas = (int)t_meas.size();
    const int tot_iter = static_cast<int>(1.0 / dt) + 1 + n_meas;                                   // R/:1832
    const double t_max = t_meas[n_meas - 1];                                                        // R/:1833
    const int W_ = 4 * (Sp - 2) + 3;
    D2Vec_Double acc(tot_iter, vector<double>(W_, 0.0));
    int ind = 0;
    for (double t = 0; t <= 1; t += dt) { if (ind < tot_iter) acc[ind][0] = t; ind += 1; }          // R/:1840-1841
    for (int i = 0; i < n_meas; i++) if (ind + i < tot_iter) acc[ind + i][0] = t_meas[i];           // R/:1842-1843

    const double beta = c * gmax - d;                                                               // R/:1851
    double p0istar, p0jstar, p0mstar;                                                               // R/:1855-1866
    if (a < 1) { p0istar = 0.0; p0jstar = a / rW; p0mstar = a / (alpha * W0); }
    else { p0jstar = d * K[1] / beta; p0istar = (c / d) * (a - rW * p0jstar); p0mstar = (a / alpha) * (p0istar + K[2]) / (p0istar + K[2] * W0); }
    const double perc = 0.015;
    const double c_high[Sp] = {p0istar + dP, p0jstar, p0mstar}, c_low[Sp] = {p0istar, p0jstar, p0mstar};   // R/:1914

    tm_params p{};
    p.model = TM_MODEL_RIETKERK_DET; p.L = g; p.dt = dt; p.dx = dx; p.a = a;
    p.rc = c; p.gmax = gmax; p.alpha = alpha; p.rW = rW; p.W0 = W0; p.d = d; p.K1 = K[1]; p.K2 = K[2];
    p.D0 = D[0]; p.D1 = D[1]; p.D2 = D[2]; p.sigma0 = 1.0;
    Ctx ctx(p, r);                                         // the r replicates advance together as one batch of trials
    const size_t n = (size_t)g * g;
    vector<double> host(3 * n * r);
    for (int j = 0; j < r; j++) init_frame_noise_off(&host[3 * n * j], g, j, perc, c_high, c_low);
    ctx.ck(tm_upload(ctx.c, host.data()), "tm_upload");

    vector<D2Vec_Double> RhoM(r, D2Vec_Double(tot_iter, vector<double>(2 * Sp, 0.0)));
    vector<tm_stats> st(r);
    double t = 0; int index = 0, iter_index = 0, pending = 0;
    while (t < t_max + dt) {                                                                        // R/:1938
        const bool window = index < n_meas && t >= t_meas[index] - dt / 2.0 && t < t_meas[index] + dt / 2.0;
        if (t <= 1.0 || window) {                                                                   // R/:1940
            if (pending) { ctx.ck(tm_step(ctx.c, pending), "tm_step"); pending = 0; }
            ctx.ck(tm_measure(ctx.c, st.data()), "tm_measure");
            for (int j = 0; j < r; j++) {
                if (iter_index >= tot_iter) break;
                for (int s = 0; s < Sp; s++) {
                    RhoM[j][iter_index][2 * s] = (double)st[j].active[s]; RhoM[j][iter_index][2 * s + 1] = st[j].mean[s];
                    if (st[j].active[s] == 0 && s == 0) break;                                      // R/:1953: the water entries stay 0
                }
            }
            iter_index += 1;
            // snapshots of every replicate at t = 0, for 500 <= t <= 2500 and in the last 15 rows (R/:1959)
            if (write_noise_off_frames && ((t >= 500 && t <= 2500) || t == 0 || (iter_index >= tot_iter - 15 && iter_index <= tot_iter))) {
                ctx.ck(tm_download(ctx.c, host.data()), "tm_download");
                for (int j = 0; j < r; j++) save_noise_off_frame(&host[3 * n * j], a, a_st, a_end, t, dt, dx, dP, D[0], D[1], alpha, gmax, j, g);
            }
            if (t > 1.0) index += 1;                                                                // R/:1987
        }
        pending += 1;
        t += dt;
    }
    for (int j = 0; j < r; j++) {                                                                   // R/:2058-2083
        const D2Vec_Double &X = RhoM[j];
        if (j == 0) {
            for (int i = 0; i < tot_iter; i++) {
                acc[i][3] = X[i][1]; acc[i][4] = 0.0; acc[i][6] = X[i][0];
                if (X[i][0] > 0) acc[i][5] = 1; else if (X[i][0] == 0) acc[i][5] = 0;
                acc[i][1] = X[i][2 * 1 + 1]; acc[i][2] = X[i][2 * 2 + 1];
            }
            continue;
        }
        for (int i = 0; i < tot_iter; i++) {                                                        // var_mean_incremental_surv_runs R/:1293
            const int rr = (int)(acc[i][5] + 1);
            if (X[i][0] == 0.0) break;
            const double mean_prev = acc[i][3];
            if (acc[i][5] == 0.0) { acc[i][3] = X[i][1]; acc[i][4] = 0.0; acc[i][6] = X[i][0]; }
            else {
                acc[i][3] = (X[i][1] + (rr - 1) * acc[i][3]) / rr;
                acc[i][6] = (X[i][0] + (rr - 1) * acc[i][6]) / rr;
                acc[i][4] = (acc[i][3] - mean_prev) * (acc[i][3] - mean_prev)
                          + (1 / (rr - 1)) * ((rr - 2) * acc[i][4] + (X[i][1] - acc[i][3]) * (X[i][1] - acc[i][3]));   // int 1/(r-1)
            }
            acc[i][5] = rr;
        }
        for (int i = 0; i < tot_iter; i++) {
            acc[i][1] = (X[i][2 * (Sp - 2) + 1] + (j)*acc[i][1]) / (j + 1);
            acc[i][2] = (X[i][2 * (Sp - 1) + 1] + (j)*acc[i][2]) / (j + 1);
        }
    }
    for (int i = 0; i < tot_iter; i++)                                                              // R/:2137-2146
        Rho.push_back({a, static_cast<double>(r), static_cast<double>(g), acc[i][0], acc[i][1], acc[i][2], acc[i][3], acc[i][4], acc[i][5], acc[i][6]});
}

// first_order_critical_exp_delta R/:2170-2345 (the noise-off sweep): a-values in turn (the reference spreads them over OpenMP
// threads and concatenates in a-order), each a batch of r replicates
void first_order_critical_exp_delta(int div, double t_max, double a_start, double a_end, double c, double gmax, double alpha, double d,
                                    double rW, double W0, double (&D)[Sp], double (&K)[3], double dt, double dx, double dP, int r, int g)
{
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = logarithm10_time_bins(t_max, dt);
    D2Vec_Double vec;
    for (size_t i = 0; i < a_space.size(); i++) {
        D2Vec_Double rows;
        RK4_Wrapper_2D(rows, t_measure, t_max, a_space[i], c, gmax, alpha, d, rW, W0, D, K, a_start, a_end, dt, dx, dP, r, g);
        vec.insert(vec.end(), rows.begin(), rows.end());
    }
    std::stringstream L, coco, tm, d3, p1, p2, rini, Dm0, Dm1, gm, alph, dix;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << a_start; p2 << a_end; rini << r; alph << alpha;
    Dm0 << std::setprecision(3) << D[0]; Dm1 << std::setprecision(3) << D[1]; gm << gmax; coco << std::setprecision(4) << c; dix << std::setprecision(2) << dx;
    const std::string dir = options().data_root + "/Data/Rietkerk/Non-Stochastic";
    mkdirs(dir);
    std::ofstream out(dir + "/1stOrderCC_Rietkerk_P_c_Delta_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_D0_" + Dm0.str() + "_D1_" + Dm1.str()
                      + "_a1_" + p1.str() + "_a2_" + p2.str() + "_dx_" + dix.str() + "_gmax_" + gm.str() + "_alpha_" + alph.str() + "_R_" + rini.str() + ".csv");   // R/:2320
    out << " a , r, L, t , <<W(x; t)>_x>_r, <<O(x; t)>_x>_r,  <<Rho0(x; t)>_x>_r, Var[<Rho0(t)>_x]_r, # Surviving Runs, # Active Sites" << "\n";   // R/:2327
    for (size_t i = 0; i < vec.size(); i++) {
        for (size_t j = 0; j + 1 < vec[i].size(); j++) out << std::setprecision(16) << vec[i][j] << ",";
        out << std::setprecision(16) << vec[i][vec[i].size() - 1] << std::endl;
    }
}

// percNoDiff_2D P/:549-662.  The steps between two logging windows run as one device batch (tm_dp_noise_off_host); the
// window test is the reference's, on the accumulated t += dt (P/:638), and `s` cycles through t_stop as there (P/:646).
void percNoDiff_2D(vector<vector<double>> &Rho, vector<double> &t_stop, double Rh0[], double t_max, double a, double b, double D,
                   double dt, double dx, int /*r*/, int g)
{
    const size_t n = (size_t)g * g;
    vector<double> Rho_dt(Rh0, Rh0 + n);
    double t = 0; int s = 0; int pending = 0;
    const int j = 0;
    auto flush = [&]() {
        if (!pending) return;
        const int rc = tm_dp_noise_off_host(current_device(), g, a, b, D, dt, dx, Rho_dt.data(), pending, nullptr);
        if (rc != 0) throw std::runtime_error(std::string("tm_dp_noise_off_host: ") + tm_last_error(nullptr));
        pending = 0;
    };
    while (t < t_max + dt) {
        pending += 1;
        t += dt;
        if (!t_stop.empty() && t > t_stop[s] - dt / 2.0 && t <= t_stop[s] + dt / 2.0) {
            flush();
            for (size_t k = 0; k < n; k++) Rho.push_back({a, static_cast<double>(j), t, static_cast<double>(k), Rho_dt[k]});
            s = (s + 1) % (int)t_stop.size();
        }
    }
    flush();
}

void rietkerk_Dornic_2D_MultiSp(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax,
                                double alpha, double rW, double W0, double (&D)[Sp], double (&)[SpB], double (&K)[3],
                                double sigma[], double a_st, double a_end, double a_c, double (&)[SpB][SpB], double (&)[SpB][SpB],
                                double (&)[SpB], double (&M)[SpB], double[], int (&)[SpB], double clow[], double dt, double dx,
                                double dP, int r, int g, double Gstar, double Vstar)
{
    vector<D2Vec_Double> per_a;
    riet_batch(per_a, vector<double>{a}, t_meas, c, gmax, alpha, rW, W0, D, K, sigma, a_st, a_end, a_c, M, clow, dt, dx, dP, r, g, Gstar, Vstar);
    Rho.insert(Rho.end(), per_a[0].begin(), per_a[0].end());
}

// Rietkerk_FPE's runner of the same vegetation-only model (F/:3697 compiled with -DSPB=1; there is nothing to move): the step is the same
// kernel pair with the clamp `v < 1e-10 && v >= 0` (F/:4460); bookkeeping in the FPE flavour (checkpoints at 85 / 90 / 95 / 100 %, its frame rule)
void rietkerk_DorFPE_2D_MultiSp(D2Vec_Double &Rho, vector<double> &t_meas, double /*t_max*/, double a, double c, double gmax,
                                double alpha, double rW, double W0, double (&D)[Sp], double (&)[SpB], double (&K)[3],
                                double sigma[], double a_st, double a_end, double a_c, double (&)[SpB][SpB], double (&)[SpB][SpB],
                                double (&)[SpB], double (&M)[SpB], double[], int (&)[SpB], double clow[], double dt, double dx,
                                double dP, int r, int g, double Gstar, double Vstar)
{
    vector<D2Vec_Double> per_a;
    riet_batch(per_a, vector<double>{a}, t_meas, c, gmax, alpha, rW, W0, D, K, sigma, a_st, a_end, a_c, M, clow, dt, dx, dP, r, g, Gstar, Vstar, true);
    Rho.insert(Rho.end(), per_a[0].begin(), per_a[0].end());
}

// The a-sweeps' two schedules, R/:4395-4459 (the same lines in every SPB build and in F/): t_measure = {0} U log10 bins (with the
// 80000..150000 linear window for long runs), returned; frame_tmeas = the same bins with [50, 400] and [400, 1200] re-gridded, set.
vector<double> sweep_time_schedule(double t_max, double dt)
{
    vector<double> t_measure = logarithm10_time_bins(t_max, 1 * dt);
    t_measure.insert(t_measure.begin(), 0.0);                                    // R/:4401
    std::sort(t_measure.begin(), t_measure.end());
    t_measure.erase(std::unique(t_measure.begin(), t_measure.end()), t_measure.end());
    frame_tmeas = t_measure;                                                       // R/:4424-4459
    frame_tmeas = switchsort_and_bait(frame_tmeas, 50, 400, 8);
    frame_tmeas = switchsort_and_bait(frame_tmeas, 400, 1200, 9);
    if (t_measure[t_measure.size() - 1] > 90000) {
        vector<double> w1 = linspace(50000, 100000, 26);
        frame_tmeas.insert(std::upper_bound(frame_tmeas.begin(), frame_tmeas.end(), w1[0]), w1.begin(), w1.end());
        std::sort(frame_tmeas.begin(), frame_tmeas.end());
        vector<double> w2 = linspace(80000, 150000, 141);
        t_measure.insert(std::upper_bound(t_measure.begin(), t_measure.end(), w2[0]), w2.begin(), w2.end());
        std::sort(t_measure.begin(), t_measure.end());
    }
    frame_tmeas.erase(std::unique(frame_tmeas.begin(), frame_tmeas.end()), frame_tmeas.end());
    t_measure.erase(std::unique(t_measure.begin(), t_measure.end()), t_measure.end());
    return t_measure;
}

namespace {
void sweep_1sp(bool fpe, int div, double t_max, double a_start, double a_end, double a_c, double c, double gmax, double alpha, double rW, double W0,
               double (&D)[Sp], double (&K)[3], double sigma[], double (&M)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
               double Gstar, double Vstar) // R/:4379-4617, F/:4665-4880
{
    vector<double> a_space = linspace(a_start, a_end, div);
    vector<double> t_measure = sweep_time_schedule(t_max, dt);

    vector<D2Vec_Double> per_a;
    riet_batch(per_a, a_space, t_measure, c, gmax, alpha, rW, W0, D, K, sigma, a_start, a_end, a_c, M, clo, dt, dx, dP, r, g, Gstar, Vstar, fpe);

    std::stringstream L, tm, d3, p1, p2, rini, Dm1, dix, geq, veq;
    L << g; tm << t_max; d3 << std::setprecision(3) << dt; p1 << a_start; p2 << a_end; rini << r;
    Dm1 << std::setprecision(3) << D[1]; dix << std::setprecision(2) << dx; geq << std::setprecision(5) << Gstar; veq << std::setprecision(5) << Vstar;
    const std::string sp = stat_prefix.empty() ? options().data_root + "/Data/Rietkerk/Stochastic/1Sp/1stCC_Rietkerk_" + prefix + "_P_c_G_" : stat_prefix;
    std::string savefile = sp + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_D1_" + Dm1.str() + "_a1_" + p1.str() + "_a2_" + p2.str() + "_dx_" + dix.str();
    if (Gstar != -1) savefile += "_Geq_" + geq.str() + "_R_" + rini.str() + ".csv";
    else if (Vstar != -1) savefile += "_Veq_" + veq.str() + "_R_" + rini.str() + ".csv";
    else savefile += "_R_" + rini.str() + ".csv";
    mkdirs(savefile.substr(0, savefile.find_last_of('/')));
    std::ofstream out(savefile);
    out << prelimheader << "\n";                                                   // R/:4601
    for (const D2Vec_Double &rows : per_a)
        for (const vector<double> &v : rows) {
            for (size_t j = 0; j + 1 < v.size(); j++) out << std::setprecision(16) << v[j] << ",";
            out << std::setprecision(16) << v[v.size() - 1] << std::endl;
        }
}
} // namespace

void first_order_critical_exp_delta_stochastic_MultiSp(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                       double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                       double (&)[SpB], double (&K)[3], double sigma[], double (&)[SpB][SpB],
                                                       double (&)[SpB][SpB], double (&)[SpB], double (&M)[SpB], double[],
                                                       int (&)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                       double Gstar, double Vstar)
{
    sweep_1sp(false, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, K, sigma, M, clo, dt, dx, dP, r, g, Gstar, Vstar);
}

void first_order_critical_exp_delta_stochastic_MultiSp_FPE(int div, double t_max, double a_start, double a_end, double a_c, double c,
                                                           double gmax, double alpha, double rW, double W0, double (&D)[Sp],
                                                           double (&)[SpB], double (&K)[3], double sigma[], double (&)[SpB][SpB],
                                                           double (&)[SpB][SpB], double (&)[SpB], double (&M)[SpB], double[],
                                                           int (&)[SpB], double clo[], double dt, double dx, double dP, int r, int g,
                                                           double Gstar, double Vstar)
{
    sweep_1sp(true, div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, K, sigma, M, clo, dt, dx, dP, r, g, Gstar, Vstar);
}


// ---- burn-in frames (INIT=2): init_csvconstframe R/:909-1015, init_exprtk_readCSVcolumns_frame R/:609-760 ------------
// Reads one frame CSV in the layout save_frame() writes ("a_c,  x,  P(x; t), W(x; t), O(x; t)": columns located by
// header name with blanks removed, as R/:665-672 does; a file without a header is taken as a_c, x, P, W, O).
// frame: 3*g*g doubles (P, W, O).  Returns the number of lattice rows read (g*g on success), -1 if the file cannot be opened,
// -2 if a named column is missing.
// n_fields = 3 (P, W, O), 4 (P, G, W, O) or 5 (P, G, Pr, W, O): the biotic columns are the first n_fields - 2 of "P(x;t)", "G(x;t)",
// "Pr(x;t)" and land in that order, water last (R/:700-705); a file without a header is taken as a_c, x, then the fields in memory order
int read_frame_csv_fields(const std::string &filename, int g, int n_fields, double *frame)
{
    if (n_fields < 3 || n_fields > 5) return -3;
    std::ifstream in(filename);
    if (!in.is_open()) { std::cerr << "Error opening file: " << filename << std::endl; return -1; }
    const size_t n = (size_t)g * g;
    static const char *biota[3] = {"P(x;t)", "G(x;t)", "Pr(x;t)"};
    const char *names[5]; int col[5];
    for (int k = 0; k < n_fields - 2; k++) names[k] = biota[k];
    names[n_fields - 2] = "W(x;t)"; names[n_fields - 1] = "O(x;t)";
    for (int k = 0; k < n_fields; k++) col[k] = 2 + k;
    std::string line;
    size_t row = 0;
    while (std::getline(in, line)) {
        if (line.find("a_c") != std::string::npos) {            // header line (R/:941)
            std::stringstream hs(line); std::string name; int idx = 0; int found[5] = {-1, -1, -1, -1, -1};
            while (std::getline(hs, name, ',')) {
                name.erase(std::remove_if(name.begin(), name.end(), [](unsigned char ch) { return std::isspace(ch); }), name.end());
                for (int k = 0; k < n_fields; k++) if (name == names[k]) found[k] = idx;
                idx++;
            }
            for (int k = 0; k < n_fields; k++) { if (found[k] < 0) { std::cerr << "Column missing in CSV file: " << names[k] << " in file " << filename << std::endl; return -2; } col[k] = found[k]; }
            continue;
        }
        if (row >= n) break;
        std::stringstream ss(line); std::string cell; int idx = 0;
        while (std::getline(ss, cell, ',')) {
            for (int k = 0; k < n_fields; k++)
                if (idx == col[k]) {
                    double v = 0;
                    try { v = std::stod(cell); } catch (const std::exception &) {
                        std::cerr << "Error: Unable to convert cell value to double. Column index: " << idx << " , for file: " << filename << std::endl;
                    }
                    frame[k * n + row] = v;
                }
            idx++;
        }
        row++;
    }
    return (int)row;
}

int read_frame_csv(const std::string &filename, int g, double *frame) { return read_frame_csv_fields(filename, g, 3, frame); }

// init_burnin_wrapper R/:836-906 for any number of fields: frames "<parendir><pattern><r'>.csv" with r' = (replicate + offset) mod (max
// replicate + 1); the reference draws the offset from mt19937_64(random_device) (R/:623-631), here it comes from options().seed.
std::function<void(double *, int, int, double, double, double, const double[])> burnin_frame_rule(const std::string &parendir,
                                                                                                   const std::string &filenamePattern, int n_fields)
{
    return [parendir, filenamePattern, n_fields](double *f, int g, int replicate, double, double, double, const double[]) {
        int maxrep = -1;
        struct stat sb;
        while (stat((parendir + filenamePattern + std::to_string(maxrep + 1) + ".csv").c_str(), &sb) == 0) maxrep++;
        if (maxrep < 0) {
            std::cerr << "RUN-TIME ERROR: no burn-in frame matches " << parendir << filenamePattern << "<r>.csv. EXITING." << std::endl;
            exit(5);                                                  // R/:886
        }
        unsigned long long s = options().seed ^ 0xB0B1B2B3ull;
        const int offset = (int)(splitmix(s) % (unsigned long long)(maxrep + 1));
        const int rr = (replicate + offset) % (maxrep + 1);
        const int got = read_frame_csv_fields(parendir + filenamePattern + std::to_string(rr) + ".csv", g, n_fields, f);
        if (got != g * g) { std::cerr << "RUN-TIME ERROR: burn-in frame has " << got << " rows, expected " << g * g << ". EXITING." << std::endl; exit(5); }
    };
}

void use_default_init_frame() { init_frame = default_init_frame; }

void use_burnin_frames(const std::string &parendir, const std::string &filenamePattern) { init_frame = burnin_frame_rule(parendir, filenamePattern, Sp); }

} // namespace tm_b200
