// drivers/order_2stoc_unity_rietkerk.cpp -- the vegetation + grazer (SPB = 2) a-sweep of Rietkerk_FPE on the B200 path: the program the
// reference's only CUDA seam lives in (rietkerk_DorFPE_2D_MultiSp with the FKE movement step).  Argv / stdin protocol, constants and
// output tree of Rietkerk_FPE/order_2stoc_unity_rietkerk.cpp: dt is fixed at the grazer's rounded advection time scale (:58-62: the first
// argument is read and ignored, :166), set_Prefix(prefix, mGrazer) puts the folders under 2Sp_mG_20 (:197), set_global_FKE_params
// (:202), attack-rate scaling (:225), scaling factors (:283), the sweep call (:329).  Initial frame: the deterministic Gaussian frame the
// runner uses as shipped (init_gaussframe, F/:3954-3962).
//   order_2stoc_unity_rietkerk dt(ignored) t_max g r a_start a_end div dP init_frac_graz aij_scale PREFIX
#include "tm_host.h"
#include "tm_host_2sp.h"
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
    namespace s2 = tm_b200::spb2;
    const double a_c = 1 / 24.0, dx = 0.1, mGrazer = 20;
    const double c = 10000, gmax = 0.05 * pow(10, -3.0) / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0;
    const double d0 = 0.00025 / 24.0, d1 = 0.01028, d2 = 0.00025 / 24.0, d3 = 0.025 / 24.0;
    const double v1 = 0.45427;
    const double dtv1 = round(0.11817455 * 100.0) / 100.0, dt = dtv1;                                   // :58-62
    double D[s2::Sp] = {d0, d1, d2, d3};
    double K[3] = {0, 5, 5000};
    double sigma[s2::Sp] = {sqrt(d0 / (dx * dx)), 1, 0, 0};
    double v[s2::SpB] = {0, v1};
    double dt_adv[s2::SpB] = {dtv1, dtv1};
    double aij = 3.6 * pow(10.0, -6.08) * pow(mGrazer, -0.37);
    const double hij = 1, ej = 0.45, m = 0.061609 * pow(mGrazer, -0.25) / 8760.0;
    double H[s2::SpB][s2::SpB] = {{0, hij}, {hij, 0.0}};
    double A[s2::SpB][s2::SpB] = {{0, aij}, {aij, 0.0}};
    double M[s2::SpB] = {d, m};
    double E[s2::SpB] = {1.0, ej};
    double pR[s2::Sp] = {0.0, 1.04285 / 2.0};

    double t_max, a_start, a_end, dP, init_frac_graz = 1.0, aij_scale = 1.0; int g, r, div; string pre;
    if (argc == 1) {
        cout << "Enter maximum duration of simulation (in hrs): "; cin >> t_max;
        cout << "Enter the desired grid-size (L) (dx = 100 m): "; cin >> g;
        cout << "Enter the desired number of replicates (r): "; cin >> r;
        cout << "Enter starting a-value: "; cin >> a_start;
        cout << "Enter ending a-value: "; cin >> a_end;
        cout << "Enter number of divisions of p (n+1): "; cin >> div;
        cout << "Enter kick for high state: "; cin >> dP;
        cout << "Enter fractional deviation from MFT for initial conditions of Grazer and Predator above R_c (1e-5 is a good choice): "; cin >> init_frac_graz;
        cout << "Enter scaling factor for grazer attacking rate (aij) which has a base value " << aij << " \t:"; cin >> aij_scale;
        cout << "Enter Prefix: "; cin >> pre;
        if (!cin) { cerr << "bad input" << endl; return 2; }
    } else if (argc == 12) {
        t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]); a_start = atof(argv[5]); a_end = atof(argv[6]);
        div = atoi(argv[7]); dP = atof(argv[8]); init_frac_graz = atof(argv[9]); aij_scale = atof(argv[10]); pre = argv[11];
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 10.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP, init_GP,  PREFIX.\n";
        return 1;
    }
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << r << endl;
    if (pre == "N/A" || pre == "n/a" || pre == "NA" || pre == "na" || pre == "N" || pre == "n" || pre == "") pre = prefix;
    s2::set_Prefix(pre, mGrazer);                                                                        // :197
    set_global_FKE_params(D, v, dt_adv, dt, dx, g, 4);                                                   // :202
    const FKEGlobals &G = fke_globals();
    for (int s = 0; s < s2::SpB; s++)
        cout << "For species " << s << " with [D, v, dtv]: " << D[s] << ", " << v[s] << ", " << dt_adv[s] << ": diffusion window " << G.sig_D_ScaledBounds[s]
             << ", advection window { " << G.sig_vD_ScaledBounds[s].first << ", " << G.sig_vD_ScaledBounds[s].second << " }, average shift "
             << G.mu_vel_prefactor[s] * v[s] << " cells\n";

    aij = aij_scale * aij; A[0][1] = aij; A[1][0] = aij;                                                 // :225
    if (aij_scale != 1.0 && aij_scale != 10 && aij_scale != 25) { cerr << "Scaling factor for aij not supported. Please use 1, 10, or 25." << endl; return 1; }
    const double Vstar = m / ((ej - m * hij) * aij);
    int dtV[s2::SpB] = {0, max(1, int(round(dtv1 / dt)))};
    double scaling_factor[2 * s2::Sp] = {5, dP / dP, 1, 1, 5, init_frac_graz, 1, 1};                     // :283 (INIT != 0)
    s2::init_frame = [](double *frame, int g_, int, double, double, double, const double clow[]) {
        init_gaussframe(frame, s2::Sp, s2::SpB, g_, clow + s2::Sp);                                      // F/:3954-3962
    };

    stringstream ast, est, veq, dPo; ast << a_start; est << a_end; dPo << dP; veq << setprecision(5) << Vstar;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str() + "_Veq_" + veq.str();
    recursive_dir_create(s2::frame_folder + tag);
    recursive_dir_create(s2::prelim_folder + tag + "/TimeSeries");
    recursive_dir_create(options().data_root + "/Data/Rietkerk/Stochastic/2Sp");
    cout << "Save directory for frames: " << s2::frame_folder << "\nSave directory for preliminary data: " << s2::prelim_folder
         << "\nSave directory for final data: " << s2::stat_prefix << "\nGPUs: " << device_list().size() << endl;
    try {
        s2::first_order_critical_exp_delta_stochastic_MultiSp_FPE(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H,
                                                                  E, M, pR, dtV, scaling_factor, dt, dx, dP, r, g, -1, Vstar);
    } catch (const exception &e) {
        cerr << "order_2stoc_unity_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
