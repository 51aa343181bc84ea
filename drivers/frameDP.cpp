// drivers/frameDP.cpp -- lattice snapshots of the first-order (cubic) DP model at a chosen time over a range of a on the B200 path: stdin
// protocol and constants of Percolation_FD/frameDP.cpp (:7-52: b = -2, c = 1, D = 6.25, sigma = 1, dx = 2.5), linked against the host
// layer: capture_frame_decay (P/:2997).
#include "tm_host.h"
#include <cmath>
#include <iostream>
#include <vector>

int main()
{
    using namespace std;
    tm_b200::independent_run_defaults();   // no TM_SEED: a fresh seed and file-name tokens per run, as the reference's random_device gives
    double a_start, a_end, b = -2, c = 1, t_max, t_event, dt; int div, r = 5, g;
    cout << "Enter desired time-step (dt): "; cin >> dt;
    cout << "Enter maximum duration of simulation (in hr): "; cin >> t_max;
    cout << "Enter time of frame capture (in hr): "; cin >> t_event;
    cout << "Enter the desired grid-size (L): "; cin >> g;
    cout << "Enter starting a-value (should be >= -b^{2}/4c): "; cin >> a_start;
    cout << "Enter ending a-value: "; cin >> a_end;
    cout << "Enter number of divisions of a (n+1): "; cin >> div;
    cout << "Enter number of replicates per value of measured a (R=5 is a good choice): "; cin >> r;
    if (!cin) { cerr << "bad input" << endl; return 2; }
    const double D = 6.25, sigma = sqrt(1.0), dx = 2.5;
    vector<double> Rho_0((size_t)g * g, 1.0);
    try {
        tm_b200::capture_frame_decay(Rho_0.data(), div, t_max, t_event, a_start, a_end, b, c, D, sigma, dt, dx, r, g);
    } catch (const exception &e) { cerr << "frameDP: " << e.what() << endl; return 1; }
    cout << "\ndone" << endl;
    return 0;
}
