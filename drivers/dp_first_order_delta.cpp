// drivers/dp_first_order_delta.cpp -- the config-1 driver on the B200 path.
// Same interactive inputs, constants and output file as Percolation_FD/1storderDP_delta.cpp
// (b=-2, c=1 :7; D=6.25, sigma=1, dx=2.5 :40; seven values read from stdin :13-33), but linked
// against the host layer (trophic_maelstorm_b200/host) instead of stochasticSPDP.cpp.
#include "tm_host.h"
#include <cmath>
#include <iostream>
#include <vector>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double a_start, a_end, b = -2, c = 1, t_max, D, sigma, dt, dx;
    int div, r, g;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter maximum duration of simulation (in hr): "; cin >> t_max;
    cout << "Enter the desired grid-size (L): "; cin >> g;
    cout << "Enter the desired number of replicates (r): "; cin >> r;
    cout << "Enter starting a-value (should be >= -b^{2}/4c): "; cin >> a_start;
    cout << "Enter ending a-value: "; cin >> a_end;
    cout << "Enter number of divisions of p (n+1): "; cin >> div;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    D = 6.25; sigma = sqrt(1.0); dx = 2.5;
    vector<double> Rho_0((size_t)g * g, 1.0);
    try {
        tm_b200::first_order_critical_exp_delta(Rho_0.data(), div, t_max, a_start, a_end, b, c, D, sigma, dt, dx, r, g);
    } catch (const exception &e) {
        cerr << "dp_first_order_delta: " << e.what() << endl;
        return 1;
    }
    cout << "\ndone" << endl;
    return 0;
}
