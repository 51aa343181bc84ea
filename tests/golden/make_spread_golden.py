"""Spreading-experiment fixtures from the UNMODIFIED reference (oracle/_ref, `make ref`):

  dp_spread.npz      spreading_exp_Dornic_2D (P/:1145) under sequential replay: 3 replicates from a single seed,
                     the runner's own output rows {thread, a, t, <N>_r, <R^2>_r, n_surv}
  ens_dp_spread.npz  the same runner with the reference's own mt19937_64 + std:: samplers, one replicate per
                     call, many calls: per-replicate N(t), R^2(t) series for the ensemble test

Run in the build container only (needs /root/reference):   python tests/golden/make_spread_golden.py
"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def seed_frame(L):
    f = np.zeros(L * L)
    f[L * (L // 2) + L // 2] = 1.0          # init_solitarytear, P/:90
    return f


def run(ref, L, params, r, t_meas, u=None):
    a, b, D, sigma, dt, dx = params
    rows = np.zeros((4096, 6)); cons = C.c_uint64(); ov = C.c_uint64()
    n = ref.ref_dp_spread_run(C.c_int(L), *[C.c_double(x) for x in (a, b, D, sigma, dt, dx)], C.c_int(r), ptr(seed_frame(L)),
                              ptr(t_meas), C.c_int(len(t_meas)), C.c_int(L // 2), C.c_int(L // 2),
                              ptr(u) if u is not None else None, C.c_uint64(0 if u is None else u.size),
                              ptr(rows), C.c_int(4096), C.byref(cons), C.byref(ov))
    assert n > 0 and ov.value == 0, (n, ov.value)
    return rows[:n].copy(), cons.value


def ens_job(args):
    L, params, t_meas, nrep = args
    os.chdir(tempfile.mkdtemp(prefix="tmspr_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_dp.so"))
    return [run(ref, L, params, 1, t_meas)[0][:, 2:6] for _ in range(nrep)]   # t, N, R2, surv


def main():
    import oracle_lib as O
    os.chdir(tempfile.mkdtemp(prefix="tmspr_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_dp.so"))
    params = (1.0548, 1.0, 0.25, 2 ** 0.5, 0.1, 1.0)     # spreading_exp_test.cpp:36, a_c of delta_frame_test.cpp
    # replay fixture
    L, r, n_u = 16, 3, 400000
    t_meas = np.array([1.5, 2.0, 3.0, 4.0])
    u = np.random.RandomState(31).random_sample(n_u)
    rows, consumed = run(ref, L, params, r, t_meas, u)
    np.savez_compressed(os.path.join(HERE, "dp_spread.npz"), L=L, params=np.array(params), r=r, t_meas=t_meas, seed=31, n_u=n_u,
                        rows=rows, consumed=consumed)
    print("wrote dp_spread", rows.shape, consumed)
    # ensemble fixture
    L = 48
    buf = np.zeros(4096)
    n = O.oracle().tmo_logarithm10_time_bins(30.0, 0.1, O.ptr(buf), 4096)
    t_meas = buf[:n].copy()
    ncpu = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(ncpu) as pool:
        res = pool.map(ens_job, [(L, params, t_meas, 40)] * ncpu)
    series = np.array([s for rr in res for s in rr])
    np.savez_compressed(os.path.join(HERE, "ens_dp_spread.npz"), L=L, params=np.array(params), t_meas=t_meas, series=series)
    print("wrote ens_dp_spread", series.shape)


if __name__ == "__main__":
    main()
