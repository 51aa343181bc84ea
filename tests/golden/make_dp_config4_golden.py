"""BASELINE config 4 at full size from the UNMODIFIED reference: the DP a-sweep near criticality at L = 2048
(`percolationDornic_2D`, Percolation_FD/stochasticSPDP.cpp:940, constants of delta_frame_test.cpp:5,40-49), one
replicate per a-value with the reference's own mt19937_64 / std:: samplers, seeded from std::random_device as shipped.

Run in the build container only (about 25 minutes on 8 cores; the 150 MB neighbour table of the runner lives on
its stack, hence the unlimited stack):

    (ulimit -s unlimited; python tests/golden/make_dp_config4_golden.py)

Writes tests/golden/ens_dp_L2048.npz: a_values[8], t_meas[], series[8][rows][3] = {t, <rho>, N_active}.  At L = 2048 a
lattice holds 4.2 M sites, so <rho>(t) of ONE replicate already has a relative standard error of ~1e-3: the GPU
test (tests/test_gpu_fullsize_properties.py::test_dp_L2048_sweep_matches_reference_ensemble) compares its own
replicates with these within the measured replicate-to-replicate spread.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

L, DT, TMAX, NA = 2048, 0.1, 100.0, 8


def main():
    from make_ensemble_golden import dp_job, time_bins
    a_values = np.linspace(1.0524, 1.0554, NA)
    tm = time_bins(TMAX, DT)
    jobs = [(0, L, (float(a), 1.0, 0.0, 0.25, 2 ** 0.5, DT, 1.0), np.ones(L * L), tm, 1) for a in a_values]
    with mp.get_context("spawn").Pool(min(NA, os.cpu_count() or 1)) as pool:
        res = pool.map(dp_job, jobs)
    series = np.array([r[0] for r in res])
    np.savez_compressed(os.path.join(HERE, "ens_dp_L2048.npz"), L=L, dt=DT, a_values=a_values, t_meas=tm, series=series,
                        params=np.array([1.0, 0.25, 2 ** 0.5, DT, 1.0]))
    print("wrote ens_dp_L2048.npz", series.shape)


if __name__ == "__main__":
    main()
