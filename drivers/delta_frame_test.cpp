// drivers/delta_frame_test.cpp -- the DP critical-decay sweep (config 4's program) on the B200 path: stdin protocol, constants
// (b = 1, D = 0.25, sigma = sqrt 2, dx = 1) and output file of Percolation_FD/delta_frame_test.cpp (:7-49), linked against the host layer:
// critical_exp_delta (P/:1309) from a full frame.
#include "tm_host.h"
#include <cmath>
#include <iostream>
#include <vector>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double a_start, a_end, b = 1, t_max, dt; int div, r, g;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter maximum duration of simulation (in hr): "; cin >> t_max;
    cout << "Enter the desired grid-size (L): "; cin >> g;
    cout << "Enter the desired number of replicates (r): "; cin >> r;
    cout << "Enter starting a-value: "; cin >> a_start;
    cout << "Enter ending a-value: "; cin >> a_end;
    cout << "Enter number of divisions of p (n+1): "; cin >> div;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    const double D = 0.25, sigma = sqrt(2.0), dx = 1.0;
    vector<double> Rho_0((size_t)g * g, 1.0);   // init_fullframe, P/:39
    try {
        tm_b200::critical_exp_delta(Rho_0.data(), div, t_max, a_start, a_end, b, D, sigma, dt, dx, r, g);
    } catch (const exception &e) { cerr << "delta_frame_test: " << e.what() << endl; return 1; }
    cout << "\ndone" << endl;
    return 0;
}
