// drivers/order_2stoc_test_rietkerk.cpp -- the vegetation + grazer (SPB = 2) Dornic a-sweep on the B200 path: the argv / stdin protocol,
// model constants, MFT initial conditions and output tree of Rietkerk_FD/order_2stoc_test_rietkerk.cpp (constants :44-94, stdin prompts
// :117-156, eleven-argument form :159-172, prefix and folders :197-203, 303-312, attack-rate scaling :208-210, dtV :217, the MFT
// expression strings :226-254, scaling factors :275-281, the sweep call :315), linked against the host layer instead of
// rietkerk_bjork_basic.cpp.
//   order_2stoc_test_rietkerk dt t_max g r a_start a_end div dP init_frac_graz aij_scale PREFIX
// Outputs go under <TM_DATA_ROOT>/Data/Rietkerk/... (default "..").  The a-values are dealt over every visible GPU.
#include "tm_host.h"
#include "tm_host_2sp.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

// the reference writes its MFT constants into expression strings with std::to_string (six decimals) and parses them back
static double six(double x) { return std::stod(std::to_string(x)); }

int main(int argc, char *argv[])
{
    using namespace std;
    using namespace tm_b200;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    namespace s2 = tm_b200::spb2;
    const double a_c = 1 / 24.0;
    const double c = 10000, gmax = 0.05 * pow(10, -3.0) / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0;   // :17
    const double dx = 0.1;
    const double d0 = 0.00025 / 24.0, d1 = 0.01028, d2 = 0.00025 / 24.0, d3 = 0.025 / 24.0;                                       // :55
    const double v1 = 0.45427, dtv1 = 0.11817455;                                                                                  // :62-65
    double D[s2::Sp] = {d0, d1, d2, d3};
    double K[3] = {0, 5, 5000};
    double sigma[s2::Sp] = {sqrt(d0 / (dx * dx)), 1, 0, 0};
    double v[s2::SpB] = {0, v1};
    double aij = 3.6 * pow(10.0, -6.08) * pow(20.0, -0.37);                                                                        // :83
    const double hij = 1, ej = 0.45, m = 0.061609 * pow(20.0, -0.25) / 8760.0;
    double H[s2::SpB][s2::SpB] = {{0, hij}, {hij, 0.0}};
    double A[s2::SpB][s2::SpB] = {{0, aij}, {aij, 0.0}};
    double M[s2::SpB] = {d, m};
    double E[s2::SpB] = {1.0, ej};
    double pR[s2::Sp] = {0.0, 1.04285 / 2.0};
    const double Wstar_veg = d * K[1] / (c * gmax - d);

    double dt, t_max, a_start, a_end, dP, init_frac_graz = 1.0, aij_scale = 1.0; int g, r, div; string pre;
    if (argc == 1) {
        cout << "Enter desired time-step (~0.1 is a good choice): "; cin >> dt;
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
        dt = atof(argv[1]); t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]); a_start = atof(argv[5]); a_end = atof(argv[6]);
        div = atoi(argv[7]); dP = atof(argv[8]); init_frac_graz = atof(argv[9]); aij_scale = atof(argv[10]); pre = argv[11];
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 10.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP, init_GP,  PREFIX.\n";   // the reference's own text, :175-177
        return 1;
    }
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << r << endl;
    if (2 * v[1] * dt / dx > 1) cout << "Warning: CFL condition violated. The velocity of the grazer is too high for the given dx and dt." << endl;
    else cout << "CFL condition satisfied." << endl;
    if (pre == "N/A" || pre == "n/a" || pre == "NA" || pre == "na" || pre == "N" || pre == "n" || pre == "") pre = prefix;
    s2::set_Prefix(pre);   // prefix and the three output locations, R/:113-135 with SpB = 2

    aij = aij_scale * aij; A[0][1] = aij; A[1][0] = aij;                                                                           // :208
    const double Vstar = m / ((ej - m * hij) * aij);
    int dtV[s2::SpB] = {0, max(1, int(round(dtv1 / dt)))};                                                                         // :217
    double A_G, b_G;                                                                                                               // :230-243
    if (aij_scale == 1.0) { A_G = 7332.541; b_G = -15.5375; }
    else if (aij_scale == 10) { A_G = 645.346; b_G = -18.94669; }
    else if (aij_scale == 25) { A_G = 240; b_G = -20.7705; }
    else { cerr << "Scaling factor for aij not supported. Please use 1, 10, or 25." << endl; return 1; }
    const double mW = 1 / rW, cW = 0.0, mO = (Vstar + K[2]) / ((Vstar + K[2] * W0) * alpha), mO_Prev = 1 / (alpha * W0);
    double scaling_factor[2 * s2::Sp] = {5, dP / dP, 1, 1, 5, init_frac_graz, 1, 1};                                               // :280 (INIT != 0)

    // MFT_Vec_CoexExpr = {V_prev, G_prev, W_prev, O_prev, V, G, W, O} (:245-254, 283) evaluated at a, constants as the strings carry them
    s2::init_frame = [=](double *frame, int g_, int replicate, double a, double a_c_, double dP_, const double clow[]) {
        const double expr[2 * s2::Sp] = {0.0, 0.0, six(Wstar_veg), six(mO_Prev) * a,
                                         six(Vstar), six(A_G) * (1 - exp(six(b_G) * (a - six(a_c)))), six(mW) * a - six(cW), six(mO) * a};
        init_randbiMFTframe(frame, s2::Sp, g_, replicate, a, a_c_, dP_, 0.015, clow, expr);                                        // perc: R/:3403
    };

    stringstream ast, est, veq, dPo; ast << a_start; est << a_end; dPo << dP; veq << setprecision(5) << Vstar;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str() + "_Veq_" + veq.str();
    recursive_dir_create(s2::frame_folder + tag);
    recursive_dir_create(s2::prelim_folder + tag + "/TimeSeries");
    recursive_dir_create(options().data_root + "/Data/Rietkerk/Stochastic/2Sp");
    cout << "Save directory for frames: " << s2::frame_folder << "\nSave directory for preliminary data: " << s2::prelim_folder
         << "\nSave directory for final data: " << s2::stat_prefix << "\nGPUs: " << device_list().size() << endl;
    try {
        s2::first_order_critical_exp_delta_stochastic_MultiSp(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E,
                                                              M, pR, dtV, scaling_factor, dt, dx, dP, r, g, -1, Vstar);
    } catch (const exception &e) {
        cerr << "order_2stoc_test_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
