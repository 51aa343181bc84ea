// drivers/order_1stoc_test_rietkerk.cpp -- the Rietkerk (vegetation / soil water / surface water, SPB = 1) Dornic a-sweep on the
// B200 path: the argv and stdin protocol, model constants and output tree of Rietkerk_FD/order_1stoc_test_rietkerk.cpp
// (constants :17-48, stdin prompts :66-98, nine-argument form :99-110, folders :163-173, the sweep call :175), linked against
// the host layer (trophic_maelstorm_b200/host) instead of rietkerk_bjork_basic.cpp.
//   order_1stoc_test_rietkerk dt t_max g r a_start a_end div dP PREFIX
// Outputs go under <TM_DATA_ROOT>/Data/Rietkerk/... (default "..", as the reference writes "../Data/Rietkerk/...").
// Sweeps are spread over every visible GPU (TM_DEVICES / TM_DEVICE_LIST / TM_DEVICE narrow that); results do not depend on it.
// Compiled with -DTM_FPE_PROGRAM (drivers/order_1stoc_test_rietkerk_fpe) it is Rietkerk_FPE/order_1stoc_test_rietkerk.cpp -- the same
// source text there, linked against rietkerk_bjork_FKE.cpp: the FPE runner (clamp at 1e-10, its checkpoint and frame rules) starting from
// the deterministic Gaussian frame that runner uses as shipped (F/:3954-3962).
#include "tm_host.h"
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
    // units: kg, km, hr (Bonachela et al. 2015), :28
    const double c = 10000, gmax = 0.05 * pow(10, -3.0) / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0;
    const double dx = 0.1, a_c = 1 / 24.0;
    double D[Sp] = {0.00025 / 24.0, 0.00025 / 24.0, 0.025 / 24.0};
    double K[3] = {0, 5, 5000};
    double sigma[Sp] = {sqrt(D[0] / (dx * dx)), 0, 0};
    double v[SpB] = {0};
    double A[SpB][SpB], H[SpB][SpB], E[SpB] = {1.0}, M[SpB] = {d}, pR[SpB] = {0.0};
    int dtV[SpB] = {0};
    const double Wstar = d * K[1] / (c * gmax - d);

    double dt, t_max, a_start, a_end, dP; int g, r, div; string pre;
    if (argc == 1) {
        cout << "Enter desired time-step (~0.25 is a good choice): "; cin >> dt;
        cout << "Enter maximum duration of simulation (in hrs): "; cin >> t_max;
        cout << "Enter the desired grid-size (L) (dx = 100 m): "; cin >> g;
        cout << "Enter the desired number of replicates (r): "; cin >> r;
        cout << "Enter starting a-value: "; cin >> a_start;
        cout << "Enter ending a-value: "; cin >> a_end;
        cout << "Enter number of divisions of p (n+1): "; cin >> div;
        cout << "Enter kick for high state: "; cin >> dP;
        cout << "Enter Prefix: "; cin >> pre;
        if (!cin) { cerr << "bad input" << endl; return 2; }
    } else if (argc == 10) {
        dt = atof(argv[1]); t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]);
        a_start = atof(argv[5]); a_end = atof(argv[6]); div = atoi(argv[7]); dP = atof(argv[8]); pre = argv[9];
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 9.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP, PREFIX.\n";
        return 1;
    }
    if (pre == "N/A" || pre == "n/a" || pre == "NA" || pre == "na" || pre == "N" || pre == "n" || pre == "") pre = prefix;
    set_Prefix(pre);
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << r << "\n"
         << "Save directory for frames: " << frame_folder << "\nSave directory for preliminary data: " << prelim_folder
         << "\nSave directory for final data: " << stat_prefix << "\nGPUs: " << device_list().size() << endl;

    double scaling_factor[2 * Sp] = {0, 1, 1, 5, 1, 1};   // :160: low state below a_c, 5 x MFT vegetation speckles above
    stringstream ast, est, dPo, geq; ast << a_start; est << a_end; dPo << dP; geq << setprecision(5) << Wstar;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str() + "_Geq_" + geq.str();
    recursive_dir_create(frame_folder + tag);
    recursive_dir_create(prelim_folder + tag + "/TimeSeries");
    recursive_dir_create(options().data_root + "/Data/Rietkerk/Stochastic/" + to_string(SpB) + "Sp");
    try {
#ifdef TM_FPE_PROGRAM
        init_frame = [](double *frame, int g_, int, double, double, double, const double clow[]) { init_gaussframe(frame, Sp, SpB, g_, clow + Sp); };
        first_order_critical_exp_delta_stochastic_MultiSp_FPE(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E,
                                                              M, pR, dtV, scaling_factor, dt, dx, dP, r, g, Wstar);
#else
        first_order_critical_exp_delta_stochastic_MultiSp(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E,
                                                          M, pR, dtV, scaling_factor, dt, dx, dP, r, g, Wstar);
#endif
    } catch (const exception &e) {
        cerr << "order_1stoc_test_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
