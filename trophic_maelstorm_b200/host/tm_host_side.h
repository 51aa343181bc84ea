// tm_host_side.h -- the replicate runners' CSV side outputs for the multi-species models (SPB >= 2), shared by tm_host_2sp.cpp and
// tm_host_3sp.cpp: what rietkerk_Dornic_2D_MultiSp / rietkerk_DorFPE_2D_MultiSp write besides the rows they return.
//   measure()        the movement averages Rho_Mov of a measurement bin (R/:3670-3700: <gamma>, Var[gamma] -- accumulated in a FLOAT, as
//                    variance_of_vector does, R/:1221-1232 --, <v_x>, <v_y> per consumer)
//   checkpoint()     save_prelimfileswrapper (R/:3029-3140): TimeSeries/PRELIM_TSERIES_* and PRELIM_MOVSERIES_* of one replicate at the
//                    60 / 85 / 90 / 95 / 100 % bins; when the vegetation is gone, the remaining checkpoints at once and "stop" to the caller
//   frames()         save_framefileswrapper (R/:2886-2925): FRAME_* of all fields and GAMMA_* of the perception fields
//   end_replicate()  the running table so far as PRELIM_AGGRAND_* (R/:4327-4349)
// File names, headers and number formats are the reference's (save_frame R/:2836, save_prelimframe R/:2720: default 6-digit precision);
// its two random file-name tokens (thrID -> "jID", R/:3634; the per-replicate ID, R/:4343) come from jid_token() / id_token() (tm_host.h): the
// constant 0 in a reproducible run (TM_SEED set), drawn from the run's seed otherwise -- as in the SPB = 1 layer.
// Also here: Stepper, the seam between a runner's bookkeeping and the device, and run_replicates(), the replicate / time loop both
// multi-species runners share (R/:3501-4372).  The product's Stepper is the device context of tm_consumer.h / tm_consumer3.h; the CPU
// tests hand run_replicates() a Stepper of their own (test_stepper in tm_host_2sp.h / tm_host_3sp.h) to pin the bookkeeping and the
// writers on the output tree of the unmodified reference (tests/golden/side_outputs_2sp.json.gz).
#pragma once
#include "tm_host.h"
#include <cmath>
#include <functional>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace tm_b200 { namespace side {

// one replicate's time stepping as the runners see it
struct Stepper {
    virtual ~Stepper() {}
    virtual void upload(const double *frame) = 0;                               // Sp fields; gamma and v_eff start as zeros
    virtual void run(uint32_t replicate, uint32_t step0, int nsteps) = 0;
    virtual void download(double *frame, double *gamma, double *veff) = 0;      // gamma [SpB-1][n], veff [SpB-1][n][2] may be null
};

struct Config {
    int Sp = 4, SpB = 2, g = 0;
    bool fpe = false;                             // Rietkerk_FPE's runner: checkpoints at 85 / 90 / 95 / 100 % only (F/:3032), its own frame rule (F/:2886, 2993)
    double a = 0, a_st = 0, a_end = 0, dt = 0, dx = 0, dP = 0, D1 = 0, D2 = 0, Gstar = -1, Vstar = -1, t_max = 0;
    std::string frame_folder, prelim_folder;      // "<...>/Frames/Stochastic/<n>Sp/<prefix>_", "<...>/Prelims/Stochastic/<n>Sp/<prefix>_"
    std::string frame_header, gammaheader, prelimheader, movmheader;
};

class Writer {
public:
    Writer(const Config &c, int tot_iter) : c_(c), tot_iter_(tot_iter) {}
    void begin_replicate() { mov_.assign(tot_iter_, std::vector<double>(4 * c_.SpB, 0.0)); }

    void measure(int index, const double *gamma, const double *veff)
    {
        const size_t n = (size_t)c_.g * c_.g;
        for (int s = 1; s < c_.SpB; s++) {
            const double *gm = gamma + (size_t)(s - 1) * n, *ve = veff + 2 * (size_t)(s - 1) * n;
            const double mean = mean_of(gm, n, 1);
            float sum = 0.0;                                                    // R/:1224: a float accumulator
            for (size_t i = 0; i < n; ++i) sum += (gm[i] - mean) * (gm[i] - mean);
            mov_[index][4 * s] = mean; mov_[index][4 * s + 1] = sum / (double)n;
            mov_[index][4 * s + 2] = mean_of(ve, n, 2); mov_[index][4 * s + 3] = mean_of(ve + 1, n, 2);
        }
    }

    // returns 1 when the vegetation is gone at this bin (the runner then leaves the replicate's time loop, R/:3714)
    int checkpoint(int index, int j, double t, const std::vector<std::vector<double>> &Rho_M, const std::vector<double> &t_meas)
    {
        write_series(index, j, t, Rho_M, t_meas);
        if (Rho_M[index][0] != 0.0) return 0;
        const int marks[5] = {int(tot_iter_ * 0.6), int(tot_iter_ * 0.85), int(tot_iter_ * 0.9), int(tot_iter_ * 0.95), int(tot_iter_ - 1)};
        for (int k = c_.fpe ? 1 : 0; k < 5; k++)                                                 // F/:3091: no 60 % mark
            if (index <= marks[k]) write_series(marks[k], j, t_meas[marks[k]], Rho_M, t_meas);   // R/:3094-3131
        return 1;
    }
    bool is_checkpoint(int index) const
    {
        return (!c_.fpe && index == int(tot_iter_ * 0.6)) || index == int(tot_iter_ * 0.85) || index == int(tot_iter_ * 0.9)
            || index == int(tot_iter_ * 0.95) || index == int(tot_iter_ - 1);                    // R/:3703, F/:3032
    }

    void frames(double t, int j, const double *frame, const double *gamma)
    {
        const size_t n = (size_t)c_.g * c_.g;
        std::stringstream L, tm, d3, p1, Dm1, dix;
        L << c_.g; tm << t; d3 << std::setprecision(3) << c_.dt; p1 << std::setprecision(4) << c_.a; dix << std::setprecision(2) << c_.dx;
        Dm1 << std::setprecision(3) << c_.D1;
        const std::string dir = c_.frame_folder + run_tag();
        const std::string tail = L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str() + "_D1_" + Dm1.str() + "_jID_" + jid_token(c_.a) + "_dx_" + dix.str() + "_R_";
        std::vector<std::vector<double>> fr(c_.Sp, std::vector<double>(n)), gm(c_.SpB, std::vector<double>(n, 0.0));
        for (int s = 0; s < c_.Sp; s++) std::copy(frame + s * n, frame + (s + 1) * n, fr[s].begin());
        for (int s = 1; s < c_.SpB; s++) std::copy(gamma + (s - 1) * n, gamma + s * n, gm[s].begin());
        save_frame(fr, dir, "/FRAME_P_c_DP_G_" + tail, c_.a, c_.a_st, c_.a_end, t, c_.dt, c_.dx, c_.dP, j, c_.g, c_.frame_header);
        save_frame(gm, dir, "/GAMMA_G_" + tail, c_.a, c_.a_st, c_.a_end, t, c_.dt, c_.dx, c_.dP, j, c_.g, c_.gammaheader);
    }

    void end_replicate(int j, std::vector<std::vector<double>> &rho_rep_avg_var)
    {
        std::stringstream L, tm, d3, p1, Dm;
        L << c_.g; tm << c_.t_max; d3 << std::setprecision(3) << c_.dt; p1 << std::setprecision(5) << c_.a; Dm << std::setprecision(4) << c_.D2;
        std::string dir = c_.prelim_folder + run_tag();
        if (c_.Gstar == -1 && c_.Vstar != -1) dir += "/TimeSeries";              // R/:4337: the reference's own slip, kept
        const std::string pattern = "/PRELIM_AGGRAND_P_c_ID_" + id_token(c_.a, j) + "_DP_G_" + L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str() + "_D2_" + Dm.str() + "_R_";
        save_prelimframe(rho_rep_avg_var, dir, pattern, c_.a, c_.a_st, c_.a_end, c_.t_max, c_.dt, c_.dx, c_.dP, j, c_.g, c_.prelimheader, true);
    }

private:
    static double mean_of(const double *x, size_t n, size_t stride)
    {
        double s = 0.0;
        for (size_t i = 0; i < n; ++i) s += x[i * stride];
        return s / (double)n;
    }
    std::string run_tag() const
    {
        std::stringstream a1, a2, dimitri, geq, veq;
        a1 << c_.a_st; a2 << c_.a_end; dimitri << c_.dP; geq << std::setprecision(5) << c_.Gstar; veq << std::setprecision(5) << c_.Vstar;
        std::string t = a1.str() + "-" + a2.str() + "_dP_" + dimitri.str();
        if (c_.Gstar != -1) t += "_Geq_" + geq.str(); else if (c_.Vstar != -1) t += "_Veq_" + veq.str();
        return t;
    }
    // var_mean_incremental_surv_runs(zeros, Rho_M, idx + 1, 0) (R/:3040) and generic_SPBmean_surv_runs(zeros, Rho_Mov, idx + 1, 4, 0) (R/:3045)
    void write_series(int idx, int j, double t, const std::vector<std::vector<double>> &X, const std::vector<double> &t_meas)
    {
        const int Sp = c_.Sp, SpB = c_.SpB;
        std::vector<std::vector<double>> tab(idx + 1, std::vector<double>(4 * SpB + 3, 0.0));
        // Rietkerk_FPE's save_prelimframe (F/:2758-2763) always walks 3 + 4 SpB columns, two past the end of a movement row (heap
        // garbage in the reference's files): the same number of columns here, zeros in the last two
        std::vector<std::vector<double>> mov(idx + 1, std::vector<double>(4 * SpB + 1 + (c_.fpe ? 2 : 0), 0.0));
        for (int s = 0; s < SpB; s++)
            for (int i = 0; i <= idx; i++) {
                if (X[i][2 * s] == 0.0) break;
                tab[i][4 * s + 3] = X[i][2 * s + 1]; tab[i][4 * s + 4] = 0.0; tab[i][4 * s + 6] = X[i][2 * s]; tab[i][4 * s + 5] = 1;
            }
        for (int i = 0; i <= idx; i++) {
            tab[i][1] = X[i][2 * (Sp - 2) + 1]; tab[i][2] = X[i][2 * (Sp - 1) + 1]; tab[i][0] = t_meas[i];
            for (int k = 0; k < 4 * SpB; k++) mov[i][k + 1] = mov_[i][k];
            mov[i][0] = t_meas[i];
        }
        std::stringstream L, tm, d3, p1;
        L << c_.g; tm << t; d3 << std::setprecision(3) << c_.dt; p1 << std::setprecision(5) << c_.a;
        const std::string dir = c_.prelim_folder + run_tag() + "/TimeSeries";
        const std::string tail = L.str() + "_T_" + tm.str() + "_dt_" + d3.str() + "_a_" + p1.str() + "_jID_" + jid_token(c_.a) + "_R_";
        save_prelimframe(tab, dir, "/PRELIM_TSERIES_P_c_DP_G_" + tail, c_.a, c_.a_st, c_.a_end, t, c_.dt, c_.dx, c_.dP, j, c_.g, c_.prelimheader, false, false);
        save_prelimframe(mov, dir, "/PRELIM_MOVSERIES_P_c_DP_G_" + tail, c_.a, c_.a_st, c_.a_end, t, c_.dt, c_.dx, c_.dP, j, c_.g, c_.movmheader, false, false);
    }
    Config c_;
    int tot_iter_;
    std::vector<std::vector<double>> mov_;
};


// What a runner hands to run_replicates besides the Stepper
struct Run {
    int r = 1;
    double a_c = 0;
    const double *clow = nullptr;
    std::function<void(double *, int, int, double, double, double, const double[])> init_frame;
    std::function<void(int, int, double, const double *)> on_measure;                              // may be empty
    std::function<void(std::vector<std::vector<double>> &, const std::vector<std::vector<double>> &, int, int)> var_mean;   // R/:1293 for this SpB
};

// The replicate loop of rietkerk_Dornic_2D_MultiSp / rietkerk_DorFPE_2D_MultiSp for SPB >= 2 (R/:3501-4372, F/:3850-4650): measurement
// schedule from the accumulated t += dt (R/:3648, 3659), Rho_M = {#occupied, <rho>} per field and the movement averages at every bin,
// checkpoint tables and the stop of a replicate whose vegetation is gone (R/:3703-3716), frames on the frame_tmeas schedule with the
// reference's save condition (R/:3736-3748), replicate-order running statistics (R/:4294-4322), the running table after every
// replicate (R/:4327-4349) and the rows handed back (R/:4361-4372).  Steps between two bookkeeping events are enqueued as one batch.
inline void run_replicates(Stepper &st, const Config &c, const Run &R, const std::vector<double> &t_meas, std::vector<std::vector<double>> &Rho)
{
    using D2 = std::vector<std::vector<double>>;
    const int Sp = c.Sp, SpB = c.SpB, tot_iter = (int)t_meas.size(), W_ = 4 * SpB + 3;
    const size_t n = (size_t)c.g * c.g;
    const double dt = c.dt, t_end = t_meas.back();                                    // R/:3439
    const std::vector<double> &ftm = frame_tmeas;
    const bool prelims = write_prelims;
    Writer w(c, tot_iter);
    D2 rho_rep_avg_var(tot_iter, std::vector<double>(W_, 0.0));
    for (int i = 0; i < tot_iter; i++) rho_rep_avg_var[i][0] = t_meas[i];            // R/:3484-3487
    std::vector<double> frame(Sp * n), gamma((SpB - 1) * n), veff(2 * (SpB - 1) * n);

    for (int j = 0; j < R.r; j++) {
        R.init_frame(frame.data(), c.g, j, c.a, R.a_c, c.dP, R.clow);
        st.upload(frame.data());
        D2 Rho_M(tot_iter, std::vector<double>(2 * Sp, 0.0));
        w.begin_replicate();
        double t = 0; int index = 0; size_t frame_index = 0; int pending = 0; uint32_t step = 0;
        bool fresh = false;                                                           // host copies equal the device state
        auto sync = [&](bool movement) {
            if (pending) { st.run((uint32_t)j, step, pending); step += pending; pending = 0; fresh = false; }
            if (!fresh) { st.download(frame.data(), movement ? gamma.data() : nullptr, movement ? veff.data() : nullptr); fresh = movement; }
        };
        while (t < t_end + dt) {                                                      // R/:3648
            const bool meas = index < tot_iter && (t == 0 || (t >= t_meas[index] - dt / 2.0 && t < t_meas[index] + dt / 2.0));       // R/:3659
            const bool snap = frame_index < ftm.size() && (t == 0 || (t >= ftm[frame_index] - dt / 2.0 && t < ftm[frame_index] + dt / 2.0));   // R/:3736
            if (meas) {
                sync(prelims || snap);
                for (int s = 0; s < Sp; s++) {
                    const double *f = &frame[s * n];
                    double sum = 0.0, occ = 0.0;                                     // mean_of_vector R/:1211, occupied_sites_of_vector R/:1245
                    for (size_t i = 0; i < n; i++) {
                        if (f[i] != f[i]) { std::cerr << "RHO NAAN PARATHA in field " << s << " at t=" << t << "\n"; exit(3); }   // R/:4156
                        sum += f[i];
                        if (f[i] > 0) occ += 1;
                    }
                    Rho_M[index][2 * s] = occ; Rho_M[index][2 * s + 1] = sum / (double)n;   // R/:3662-3668
                }
                if (prelims) w.measure(index, gamma.data(), veff.data());            // R/:3670-3700
                if (R.on_measure) R.on_measure(j, index, t, frame.data());
                if (w.is_checkpoint(index)) {                                         // R/:3703-3716, F/:4068-4072
                    if (prelims) w.checkpoint(index, j, t, Rho_M, t_meas);
                    if (Rho_M[index][0] == 0.0) break;                                // all vegetation dead: on to the next replicate
                }
                index += 1;
            }
            if (snap) {
                bool save = frame_index + 10 >= ftm.size() || (t >= 50000 && t <= 150000) || (t >= (c.fpe ? 50 : 20) && t <= 15000) || t == 0;   // R/:3739, F/:2886
                if (c.fpe && t > 1500 && t < 6000 && frame_index % 2 != 0) save = false;     // F/:2993-3004: every second frame in that window
                if (save) {
                    sync(true);
                    w.frames(t, j, frame.data(), gamma.data());
                }
                frame_index += 1;
            }
            pending += 1;
            t += dt;
        }
        if (j == 0) {                                                                 // R/:4294-4316
            for (int i = 0; i < tot_iter; i++) {
                for (int s = 0; s < SpB; s++) {
                    rho_rep_avg_var[i][4 * s + 3] = Rho_M[i][2 * s + 1]; rho_rep_avg_var[i][4 * s + 4] = 0.0;
                    rho_rep_avg_var[i][4 * s + 6] = Rho_M[i][2 * s];
                    if (Rho_M[i][2 * s] > 0) rho_rep_avg_var[i][4 * s + 5] = 1; else if (Rho_M[i][2 * s] == 0) rho_rep_avg_var[i][4 * s + 5] = 0;
                }
                rho_rep_avg_var[i][1] = Rho_M[i][2 * (Sp - 2) + 1]; rho_rep_avg_var[i][2] = Rho_M[i][2 * (Sp - 1) + 1];
            }
        } else R.var_mean(rho_rep_avg_var, Rho_M, tot_iter, j);                       // R/:4321
        if (prelims) w.end_replicate(j, rho_rep_avg_var);                             // R/:4327-4349
    }
    for (int i = 0; i < tot_iter; i++) {                                              // R/:4361-4372
        Rho.push_back({c.a, static_cast<double>(R.r), static_cast<double>(c.g), rho_rep_avg_var[i][0], rho_rep_avg_var[i][1], rho_rep_avg_var[i][2]});
        for (int s = 0; s < Sp - 2; s++)
            Rho.back().insert(Rho.back().end(), {rho_rep_avg_var[i][4 * s + 3], rho_rep_avg_var[i][4 * s + 4],
                                                 rho_rep_avg_var[i][4 * s + 5], rho_rep_avg_var[i][4 * s + 6]});
    }
}

}} // namespace tm_b200::side
