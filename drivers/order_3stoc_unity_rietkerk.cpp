// drivers/order_3stoc_unity_rietkerk.cpp -- the vegetation + grazer + predator (SPB = 3) a-sweep of Rietkerk_FPE on the B200 path
// (rietkerk_DorFPE_2D_MultiSp: perception fields with per-site advection directions, vegetation Dornic, five-field RK4, FKE movement of
// both consumers).  Argv / stdin protocol, constants and output tree of Rietkerk_FPE/order_3stoc_unity_rietkerk.cpp: dx = 0.1 km, the two
// advection time scales rounded to 0.12 and 0.77 hr and dt fixed at the smaller one whatever the first argument says (:84-85, 81 of the
// stripped listing), dt_adv = {dt, dtv1, dtv2}, set_Prefix(prefix, mGrazer, mPredator), set_global_FKE_params, the attack-rate scalings
// and their admissible values, scaling factors {500, 10 G*, 40, 4, 20, 500, 10 G*, 40, 4, 10}; the initial frame is the deterministic
// Gaussian frame the runner uses as shipped (init_gaussframe F/:3954-3962 with amp = the second half of the scaling factors).
//   order_3stoc_unity_rietkerk dt(ignored) t_max g r a_start a_end div dP init_frac_pred aij_scale ajm_scale PREFIX
#include "tm_host.h"
#include "tm_host_3sp.h"
#include "tm_host_fke.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

int main(int argc, char *argv[])
{
    using namespace std;
    using namespace tm_b200;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    namespace s3 = tm_b200::spb3;
    const double a_c = 1 / 24.0, dx = 0.1, mGrazer = 20, mPredator = 100;
    const double c = 10000, gmax = 0.05 * pow(10, -3.0) / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0;
    const double d0 = 0.00025 / 24.0, d1 = 0.0298, d2 = 0.05221, d3 = 0.00025 / 24.0, d4 = 0.025 / 24.0;
    const double v1 = 0.45427, v2 = 0.40587;
    const double dtv1 = round(0.11817455 * 100.0) / 100.0, dtv2 = round(0.765270868 * 100.0) / 100.0, dt = min(dtv1, dtv2);
    double D[s3::Sp] = {d0, d1, d2, d3, d4};
    double K[3] = {0, 5, 5000};
    double sigma[s3::Sp] = {sqrt(d0 / (dx * dx)), 1, 3, 0, 0};
    double v[s3::SpB] = {0, v1, v2};
    double dt_adv[s3::SpB] = {dt, dtv1, dtv2};
    double aij = 3.6 * pow(10.0, -6.08) * pow(mGrazer, -0.37), ajm = 3.6 * pow(10.0, -6.08) * pow(mPredator, -0.37);
    const double hij = 1, ej = 0.45, mj = 0.061609 * pow(mGrazer, -0.25) / 8760.0;
    const double hjm = 1, em = 0.85, mm = 0.061609 * pow(mPredator, -0.25) / 8760.0;
    double H[s3::SpB][s3::SpB] = {{0, hij, 0}, {hij, 0.0, hjm}, {0, hjm, 0}};
    double A[s3::SpB][s3::SpB] = {{0, aij, 0}, {aij, 0.0, ajm}, {0, ajm, 0}};
    double M[s3::SpB] = {d, mj, mm};
    double E[s3::SpB] = {1.0, ej, em};
    double pR[s3::Sp] = {0.0, 1.04285 / 2.0, 1.25536 / 2.0};

    double t_max, a_start, a_end, dP, init_frac_pred = 1.0, aij_scale = 1.0, ajm_scale = 1.0; int g, r, div; string pre;
    if (argc == 1) {
        cout << "Enter maximum duration of simulation (in hrs): "; cin >> t_max;
        cout << "Enter the desired grid-size (L) (dx = 100 m): "; cin >> g;
        cout << "Enter the desired number of replicates (r): "; cin >> r;
        cout << "Enter starting a-value: "; cin >> a_start;
        cout << "Enter ending a-value: "; cin >> a_end;
        cout << "Enter number of divisions of p (n+1): "; cin >> div;
        cout << "Enter kick for high state: "; cin >> dP;
        cout << "Enter fractional deviation from MFT for initial conditions of the Predator above R_c: "; cin >> init_frac_pred;
        cout << "Enter scaling factor for grazer attacking rate (aij) which has a base value " << aij << " \t:"; cin >> aij_scale;
        cout << "Enter scaling factor for predator attacking rate (ajm) which has a base value " << ajm << " \t:"; cin >> ajm_scale;
        cout << "Enter Prefix: "; cin >> pre;
        if (!cin) { cerr << "bad input" << endl; return 2; }
    } else if (argc == 13) {
        t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]); a_start = atof(argv[5]); a_end = atof(argv[6]);
        div = atoi(argv[7]); dP = atof(argv[8]); init_frac_pred = atof(argv[9]); aij_scale = atof(argv[10]); ajm_scale = atof(argv[11]); pre = argv[12];
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 13.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP, init_GP, PREFIX.\n";
        return 1;
    }
    (void)init_frac_pred;   // only the INIT == 0 scaling factors read it (the Gaussian frame does not)
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << r << endl;
    if (pre == "N/A" || pre == "n/a" || pre == "NA" || pre == "na" || pre == "N" || pre == "n" || pre == "") pre = prefix;
    s3::set_Prefix(pre, mGrazer, mPredator);
    set_global_FKE_params(D, v, dt_adv, dt, dx, g, 4);
    const FKEGlobals &G = fke_globals();
    for (int s = 0; s < s3::SpB; s++)
        cout << "For species " << s << " with [D, v, dtv]: " << D[s] << ", " << v[s] << ", " << dt_adv[s] << ": diffusion window " << G.sig_D_ScaledBounds[s]
             << ", advection window { " << G.sig_vD_ScaledBounds[s].first << ", " << G.sig_vD_ScaledBounds[s].second << " }, average shift "
             << G.mu_vel_prefactor[s] * v[s] << " cells\n";

    aij = aij_scale * aij; A[0][1] = aij; A[1][0] = aij;
    ajm = ajm_scale * ajm; A[1][2] = ajm; A[2][1] = ajm;
    const double Gstar = mm / ((em - mm * hjm) * ajm);
    int dtV[s3::SpB] = {0, max(1, int(round(dtv1 / dt))), max(1, int(round(dtv2 / dt)))};
    // the reference's MFT tables exist for these scalings only and it exits for any other (the tables feed the expression strings, which
    // the Gaussian frame does not read)
    auto one_of = [](double x, std::initializer_list<double> set) { return std::find(set.begin(), set.end(), x) != set.end(); };
    const bool ok = (aij_scale == 1.0 && one_of(ajm_scale, {1.0, 10.0, 25.0})) || (aij_scale == 10.0 && one_of(ajm_scale, {1.0, 10.0, 25.0}))
                 || (aij_scale == 25.0 && one_of(ajm_scale, {1.0, 10.0, 25.0, 50.0, 100.0, 150.0}));
    if (!ok) { cerr << "Scaling factor for aij / ajm not supported. Please use aij in {1, 10, 25} and ajm in {1, 10, 25} (aij = 25: also 50, 100, 150)." << endl; return 1; }
    double scaling_factor[2 * s3::Sp] = {500, Gstar * 10, 40, 4, 20, 500, Gstar * 10, 40, 4, 10};   // INIT != 0
    s3::init_frame = [](double *frame, int g_, int, double, double, double, const double clow[]) {
        init_gaussframe(frame, s3::Sp, s3::SpB, g_, clow + s3::Sp);                                // F/:3954-3962
    };

    stringstream ast, est, dPo, geq; ast << a_start; est << a_end; dPo << dP; geq << setprecision(5) << Gstar;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str() + "_Geq_" + geq.str();
    recursive_dir_create(s3::frame_folder + tag);
    recursive_dir_create(s3::prelim_folder + tag + "/TimeSeries");
    recursive_dir_create(options().data_root + "/Data/Rietkerk/Stochastic/3Sp");
    cout << "Save directory for frames: " << s3::frame_folder << "\nSave directory for preliminary data: " << s3::prelim_folder
         << "\nSave directory for final data: " << s3::stat_prefix << "\nGPUs: " << device_list().size() << endl;
    try {
        s3::first_order_critical_exp_delta_stochastic_MultiSp_FPE(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H,
                                                                  E, M, pR, dtV, scaling_factor, dt, dx, dP, r, g, Gstar);
    } catch (const exception &e) {
        cerr << "order_3stoc_unity_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
