// drivers/finite_scaling_delta.cpp -- DP finite-size scaling in the stationary window on the B200 path: stdin protocol and constants of
// Percolation_FD/finite_scaling_delta.cpp (:7-49: b = 1, D = 0.25, sigma = sqrt 2, dx = 1, t_lim = 0), linked against the host layer:
// critical_exp_finite_scaling_stationary (P/:1538).
#include "tm_host.h"
#include <cmath>
#include <iostream>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double a_c, b = 1, t_max, t_lim = 0, dt; int div, r, g_st, g_end;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter maximum duration of simulation (end of time-window) (in hr): "; cin >> t_max;
    cout << "Enter starting grid-size (L1): "; cin >> g_st;
    cout << "Enter ending grid-size (L2): "; cin >> g_end;
    cout << "Enter number of divisions of grid-sizes (n+1): "; cin >> div;
    cout << "Enter the desired number of infra-replicates (r): "; cin >> r;
    cout << "Enter critical threshold: "; cin >> a_c;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    const double D = 0.25, sigma = sqrt(2.0), dx = 1.0;
    try {
        tm_b200::critical_exp_finite_scaling_stationary(div, g_st, g_end, t_lim, t_max, a_c, b, D, sigma, dt, dx, r);
    } catch (const exception &e) { cerr << "finite_scaling_delta: " << e.what() << endl; return 1; }
    cout << "\ndone" << endl;
    return 0;
}
