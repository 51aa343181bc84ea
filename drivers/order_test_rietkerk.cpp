// drivers/order_test_rietkerk.cpp -- the NOISE-OFF Rietkerk a-sweep (f_2D / RK4_Integrate / RK4_Wrapper_2D, no sampling)
// on the B200 path: the argv and stdin protocol, model constants and output places of Rietkerk_FD/order_test_rietkerk.cpp
// (constants :12-24, stdin prompts :36-63, eight-argument form :65-76, folders :123-168, the sweep call :182), linked against the
// host layer (trophic_maelstorm_b200/host) instead of rietkerk_bjork_basic.cpp.
//   order_test_rietkerk dt t_max g r a_start a_end div dP
// Rietkerk_FPE/order_test_rietkerk.cpp is the same program in its stdin-only form (same constants, prompts and call): run this
// binary without arguments for it.  Outputs go under <TM_DATA_ROOT>/Data/Rietkerk/... (default "..", as the reference writes
// "../Data/Rietkerk/...").  The a-values are spread over every visible GPU; results do not depend on it.
#include "tm_host.h"
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

int main(int argc, char *argv[])
{
    using namespace std;
    using namespace tm_b200;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    // g, m, day (Bonachela et al. 2015), :12-24
    const double a_note = 0.575;
    const double c = 10, d = 0.25, gmax = 0.05, alpha = 0.2, W0 = 0.2, rW = 0.2;
    const double dx = 0.2;
    double D[Sp] = {0.001, 0.001, 0.1};
    double K[3] = {0, 5, 5};

    double dt, t_max, a_start, a_end, dP, r; int g, div;      // `r` is a double in the reference's main (:9), truncated at the call
    if (argc == 1) {
        cout << "Enter desired time-step: "; cin >> dt;
        cout << "Enter maximum duration of simulation (in days): "; cin >> t_max;
        cout << "Enter the desired grid-size (L) (dx = 0.2 m): "; cin >> g;
        cout << "Enter the desired number of replicates (r): "; cin >> r;
        cout << "Note that a relevant value of 'a' may be construed as:\t" << setprecision(10) << a_note << endl;
        cout << "Enter starting a-value: "; cin >> a_start;
        cout << "Enter ending a-value: "; cin >> a_end;
        cout << "Enter number of divisions of p (n+1): "; cin >> div;
        cout << "Enter kick for high state: "; cin >> dP;
        if (!cin) { cerr << "bad input" << endl; return 2; }
    } else if (argc == 9) {
        dt = atof(argv[1]); t_max = atof(argv[2]); g = atoi(argv[3]); r = atoi(argv[4]);
        a_start = atof(argv[5]); a_end = atof(argv[6]); div = atoi(argv[7]); dP = atof(argv[8]);
    } else {
        cout << "Please enter the correct number of arguments.\nThe correct number of arguments is 8.\n"
                "The arguments are as follows: dt, t_max, g, r, a_start, a_end, div, dP\n";
        return 1;
    }
    cout << "dx = " << dx << "\t dt = " << dt << "\t t_max = " << t_max << "\t g = " << g << "\t r = " << (int)r
         << "\nGPUs: " << device_list().size() << endl;

    stringstream ast, est, dPo; ast << a_start; est << a_end; dPo << dP;
    const string tag = ast.str() + "-" + est.str() + "_dP_" + dPo.str();
    const string root = options().data_root + "/Data/Rietkerk/";
    recursive_dir_create(root + "Frames/Non-Stochastic/" + tag);      // :126-127
    recursive_dir_create(root + "Prelims/Non-Stochastic/" + tag);     // :147-148
    recursive_dir_create(root + "Non-Stochastic");                    // :166-167
    try {
        first_order_critical_exp_delta(div, t_max, a_start, a_end, c, gmax, alpha, d, rW, W0, D, K, dt, dx, dP, (int)r, g);   // :182
    } catch (const exception &e) {
        cerr << "order_test_rietkerk: " << e.what() << endl;
        return 1;
    }
    return 0;
}
