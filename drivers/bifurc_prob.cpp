// drivers/bifurc_prob.cpp -- terminal densities of the first-order (cubic) DP model over a range of initial densities on the B200 path:
// stdin protocol and constants of Percolation_FD/bifurc_prob.cpp (:7-50: b = -2, c = 1, D = 6.25, sigma = 1, dx = 2.5), linked against
// the host layer: first_order_terminal_diff (P/:2548).
#include "tm_host.h"
#include <cmath>
#include <iostream>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double p_start, p_end, a, b = -2, c = 1, t_event, dt; int div, r, g;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter time of frame capture (in hr): "; cin >> t_event;
    cout << "Enter the desired grid-size (L): "; cin >> g;
    cout << "Enter a-value: "; cin >> a;
    cout << "Enter starting initial-value (should be >= 0): "; cin >> p_start;
    cout << "Enter ending initial-value (consider setting less < 2): "; cin >> p_end;
    cout << "Enter number of divisions of p (n+1): "; cin >> div;
    cout << "Enter number of replicates "; cin >> r;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    const double D = 6.25, sigma = sqrt(1.0), dx = 2.5;
    try {
        tm_b200::first_order_terminal_diff(div, t_event, p_start, p_end, a, b, c, D, sigma, dt, dx, r, g);
    } catch (const exception &e) { cerr << "bifurc_prob: " << e.what() << endl; return 1; }
    cout << "\ndone" << endl;
    return 0;
}
