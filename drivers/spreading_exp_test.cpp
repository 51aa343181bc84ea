// drivers/spreading_exp_test.cpp -- DP spreading experiments (theta, delta, z from a single seed at the centre) on the B200 path: stdin
// protocol and constants of Percolation_FD/spreading_exp_test.cpp (:7-47: b = 1, D = 0.25, sigma = sqrt 2, dx = 1, init_solitarytear),
// linked against the host layer: critical_exp_theta_delta (P/:1667).
#include "tm_host.h"
#include <cmath>
#include <iostream>
#include <vector>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double a_c, b = 1, t_max, dt; int div, r, g;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter maximum duration of simulation (in hr): "; cin >> t_max;
    cout << "Enter the desired grid-size (L): "; cin >> g;
    cout << "Enter the desired number of infra-replicates (r): "; cin >> r;
    cout << "Enter number of divisions of supra-replicates (n+1): "; cin >> div;
    cout << "Enter critical threshold: "; cin >> a_c;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    const double D = 0.25, sigma = sqrt(2.0), dx = 1.0;
    vector<double> Rho_0((size_t)g * g);
    tm_b200::init_solitarytear(Rho_0.data(), g);   // P/:90
    try {
        tm_b200::critical_exp_theta_delta(Rho_0.data(), div, t_max, a_c, b, D, sigma, dt, dx, r, g);
    } catch (const exception &e) { cerr << "spreading_exp_test: " << e.what() << endl; return 1; }
    cout << "\ndone" << endl;
    return 0;
}
