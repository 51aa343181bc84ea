// drivers/order_3stoc_test_rietkerk.cpp -- the vegetation + grazer + predator (SPB = 3) Dornic a-sweep on the B200 path: the argv / stdin
// protocol, model constants, MFT initial conditions and output tree of Rietkerk_FD/order_3stoc_test_rietkerk.cpp (dx = 1 km :13, masses
// :18, constants :24-90, interaction matrices :108-132, twelve-argument form :182-196, CFL exits :211-220, set_Prefix(prefix, mGrazer,
// mPredator) :233, attack-rate scalings :241-243, dtV :249, the MFT tables :256-335 and expression strings :337-356, scaling factors :365,
// the sweep call), linked against the host layer instead of rietkerk_bjork_basic.cpp.
//   order_3stoc_test_rietkerk dt t_max g r a_start a_end div dP init_frac_grazpred aij_scale ajm_scale PREFIX
#include "tm_host.h"
#include "tm_host_3sp.h"
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
    namespace s3 = tm_b200::spb3;
    const double a_c = 1 / 24.0, dx = 1, mGrazer = 20, mPredator = 100;
    const double c = 10000, gmax = 0.05 * pow(10, -3.0) / 24.0, d = 0.25 / 24.0, alpha = 0.2 / 24.0, W0 = 0.2, rW = 0.2 / 24.0;
    const double d0 = 0.00025 / 24.0, d1 = 0.0298, d2 = 0.05221, d3 = 0.00025 / 24.0, d4 = 0.025 / 24.0;      // :66
    const double v1 = 0.45427, v2 = 0.40587, dtv1 = 0.11817455, dtv2 = 0.765270868;                         // :76-81
    double D[s3::Sp] = {d0, d1, d2, d3, d4};
    double K[3] = {0, 5, 5000};
    double sigma[s3::Sp] = {sqrt(d0 / (dx * dx)), 1, 3, 0, 0};
    double v[s3::SpB] = {0, v1, v2};
    double aij = 3.6 * pow(10.0, -6.08) * pow(mGrazer, -0.37), ajm = 3.6 * pow(10.0, -6.08) * pow(mPredator, -0.37);
    const double hij = 1, ej = 0.45, mj = 0.061609 * pow(mGrazer, -0.25) / 8760.0;
    const double hjm = 1, em = 0.85, mm = 0.061609 * pow(mPredator, -0.25) / 8760.0;
    double H[s3::SpB][s3::SpB] = {{0, hij, 0}, {hij, 0.0, hjm}, {0, hjm, 0}};
    double A[s3::SpB][s3::SpB] = {{0, aij, 0}, {aij, 0.0, ajm}, {0, ajm, 0}};
    double M[s3::SpB] = {d, mj, mm};
    double E[s3::SpB] = {1.0, ej, em};
    double pR[s3::Sp] = {0.0, 1.04285 / 2.0, 1.25536 / 2.0};

    double dt, t_max, a_start, a_end, dP, init_frac = 1.0, aij_scale = 1.0, ajm_scale = 1.0; int g, r, div; string pre;
    if (argc == 1) {
        cout << "Enter desired time-step (~0.1 is a good choice): "; cin >> dt;
        cout << "Enter maximum duration of simulation (in hrs): "; cin >> t_max;
        cout << "Enter the desired grid-size (L) (dx = 1 km): "; cin >> g;
        cout << "Enter the desired number of replicates (r): "; cin >> r;
        cout << "Enter starting a-value: "; cin >> a_start;
        cout << "Enter ending a-value: "; cin >> a_end;
        cout << "Enter number of divisions of p (n+1): "; cin >> div;
        cout << "Enter kick for high state: "; cin >> dP;
        cout << "Enter fractional deviation from MFT for initial conditions of Grazer and Predator above R_c (1e-5 is a good choice): "; cin >> init_frac;
        cout << "Enter scaling factor for grazer attacking rate (aij) which has a base value " << aij << " \t:"; cin >> aij_scale;
        cout << "Enter scaling factor for predator attacking rate (ajm) which has a base value " << ajm << " \t:"; cin >> ajm_scale;
        cout << "Enter Prefix: "; cin >> pre;
        if (!cin) { cerr << "bad input" << endl; return 2; }
    } else if (argc == 13) {
        dt = atof(argv[1]); t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]); a_start = atof(argv[5]); a_end = atof(argv[6]);
        div = atoi(argv[7]); dP = atof(argv[8]); init_frac = atof(argv[9]); aij_scale = atof(argv[10]); ajm_scale = atof(argv[11]); pre = argv[12];
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 13.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP, init_GP, PREFIX.\n";   // the reference's own text, :200-202
        return 1;
    }
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << r << endl;
    if (2 * v[1] * dt / dx > 1) { cout << "Warning: CFL condition violated. The velocity of the grazer is too high for the given dx and dt." << endl; return 2; }
    if (2 * v[2] * dt / dx > 1) { cout << "Warning: CFL condition violated. The velocity of the predator is too high for the given dx and dt." << endl; return 2; }
    cout << "CFL condition satisfied." << endl;
    if (pre == "N/A" || pre == "n/a" || pre == "NA" || pre == "na" || pre == "N" || pre == "n" || pre == "") pre = prefix;
    s3::set_Prefix(pre, mGrazer, mPredator);                                                                 // :233

    aij = aij_scale * aij; A[0][1] = aij; A[1][0] = aij;                                                     // :241
    ajm = ajm_scale * ajm; A[1][2] = ajm; A[2][1] = ajm;
    const double Gstar = mm / ((em - mm * hjm) * ajm);
    int dtV[s3::SpB] = {0, max(1, int(round(dtv1 / dt))), max(1, int(round(dtv2 / dt)))};                    // :249

    // the MFT tables, :256-335
    const double mV = 960000.0, cV = -40000.0, mW_Prev = 1 / rW, mO = 120.0, cO = 0.0, mO_Prev = 1 / (alpha * W0);
    double mW = -3.90344309, cW = 5.18, A_W = 0, b_W = 0, A_Pr = 4.49723728e+05, b_Pr = -1.73381257;
    auto bad = [](const char *what, const char *set) { cerr << "Scaling factor for " << what << " not supported. Please use " << set << "." << endl; return 1; };
    if (aij_scale == 1.0) {
        if (ajm_scale == 1.0) { A_Pr = 4.49723728e+05; b_Pr = -1.73381257; mW = -3.90344309; cW = 5.18; }
        else if (ajm_scale == 10.0) { A_Pr = 4.55158720e+04; b_Pr = -1.70802424; mW = -0.374252; cW = 5.01475; }
        else if (ajm_scale == 25.0) { A_Pr = 2.93592196e+05; b_Pr = -0.0911713457; mW = -0.149703; cW = 5.0059; }
        else return bad("ajm", "1, 10, or 25");
    } else if (aij_scale == 25.0) {
        if (ajm_scale == 1.0) { A_Pr = 6.8933839e+05; b_Pr = -20.0807587; A_W = 1; b_W = 8.59213; mW = -32.0623; cW = 0.0243051; }
        else if (ajm_scale == 10.0) { A_Pr = 6.8798713e+04; b_Pr = -20.4693606; A_W = 4.65526; b_W = 14.9438; mW = -55.1092; cW = -0.0286443; }
        else if (ajm_scale == 25.0) { A_Pr = 2.75159916e+04; b_Pr = -20.4953090; A_W = 4.86169; b_W = 12.9258; mW = -53.9559; cW = -0.044096; }
        else if (ajm_scale == 50.0) { A_Pr = 1.37574163e+04; b_Pr = -20.5039489; A_W = 4.92608802; b_W = 3.68792022; mW = -33.60729439; cW = -0.07048515; }
        else if (ajm_scale == 100.0) { A_Pr = 6.87856303e+03; b_Pr = -20.5082747; A_W = 4.96537423; b_W = 3.19028452; mW = -53.4178045; cW = -0.0445609397; }
        else if (ajm_scale == 150.0) { A_Pr = 4.58567622e+03; b_Pr = -20.5097222; A_W = 4.97691262; b_W = 2.68932522; mW = -53.9559; cW = -0.044096; }
        else return bad("ajm", "1, 10, 25, 50, 100 or 150");
    } else if (aij_scale == 10.0) {
        if (ajm_scale == 1.0) { A_Pr = 6.11008374e+05; b_Pr = -10.8608332; mW = -16.3029; cW = 5.64117; }
        else if (ajm_scale == 10.0) { A_Pr = 6.0987417e+04; b_Pr = -10.9556230; A_W = 4.6581; b_W = 13.2612; mW = -11.889; cW = 0.263293; }
        else if (ajm_scale == 25.0) { A_Pr = 2.4392e+04; b_Pr = -10.9619; A_W = 4.8; b_W = 2.68932522; mW = -53.3597073; cW = -0.0490352778; }
        else return bad("ajm", "1, 10, or 25");
    } else return bad("aij", "1, 10, or 25");
    double scaling_factor[2 * s3::Sp] = {0, dP / dP, dP / (10.0 * dP), 1, 1, 5, init_frac, init_frac, 1, 1};   // :365 (INIT != 0)

    // MFT_Vec_CoexExpr = {V_prev, G_prev, Pr_prev, W_prev, O_prev, V, G, Pr, W, O} (:337-356) evaluated at a, constants as the strings carry them
    s3::init_frame = [=](double *frame, int g_, int replicate, double a, double a_c_, double dP_, const double clow[]) {
        const double W = A_W == 0 ? six(mW) * a + six(cW) : six(A_W) + six(b_W) * exp(six(mW) * (a - six(cW)));
        const double expr[2 * s3::Sp] = {six(0.0) * a + six(0.0), 0.0, 0.0, six(mW_Prev) * a, six(mO_Prev) * a,
                                         six(mV) * a + six(cV), six(Gstar), six(A_Pr) * (1 - exp(six(b_Pr) * (a - six(a_c)))), W, six(mO) * a + six(cO)};
        init_randbiMFTframe(frame, s3::Sp, g_, replicate, a, a_c_, dP_, 0.015, clow, expr);
    };

    stringstream ast, est, dPo, geq; ast << a_start; est << a_end; dPo << dP; geq << setprecision(5) << Gstar;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str() + "_Geq_" + geq.str();
    recursive_dir_create(s3::frame_folder + tag);
    recursive_dir_create(s3::prelim_folder + tag + "/TimeSeries");
    recursive_dir_create(options().data_root + "/Data/Rietkerk/Stochastic/3Sp");
    cout << "Save directory for frames: " << s3::frame_folder << "\nSave directory for preliminary data: " << s3::prelim_folder
         << "\nSave directory for final data: " << s3::stat_prefix << "\nGPUs: " << device_list().size() << endl;
    try {
        s3::first_order_critical_exp_delta_stochastic_MultiSp(div, t_max, a_start, a_end, a_c, c, gmax, alpha, rW, W0, D, v, K, sigma, A, H, E,
                                                              M, pR, dtV, scaling_factor, dt, dx, dP, r, g, Gstar);
    } catch (const exception &e) {
        cerr << "order_3stoc_test_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
