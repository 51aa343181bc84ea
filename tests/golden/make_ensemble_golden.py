"""Ensemble statistics of the UNMODIFIED reference (oracle/_ref, its own mt19937_64 + std:: samplers,
seeded from std::random_device as shipped): per-replicate <rho>(t) and N_active(t) series.

Run in the build container only:   python tests/golden/make_ensemble_golden.py
The GPU tests compare the product's ensembles against these within confidence intervals
(north_star: "ensemble rho(t), survival probability and DP decay exponents agree within 95 % CIs").
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
sys.path.insert(0, HERE)
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def time_bins(t_max, dt):
    import oracle_lib as O
    buf = np.zeros(4096)
    n = O.oracle().tmo_logarithm10_time_bins(t_max, dt, O.ptr(buf), 4096)
    return buf[:n].copy()


def dp_job(args):
    cubic, L, params, rho0, t_meas, nrep = args
    os.chdir(tempfile.mkdtemp(prefix="tmens_"))
    os.makedirs("../Data/SPDP", exist_ok=True)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_dp.so"))
    out = []
    for _ in range(nrep):
        rows = np.zeros((4096, 6))
        n = ref.ref_dp_run(C.c_int(cubic), C.c_int(L), *[C.c_double(x) for x in params], C.c_int(1), ptr(rho0), ptr(t_meas),
                           C.c_int(len(t_meas)), None, C.c_uint64(0), ptr(rows), C.c_int(4096), None, None)
        out.append(rows[:n, [0, 1, 4]].copy())  # t, mean, active
    return out


def riet_job(args):
    L, a, dt, init, t_meas, nrep = args
    from make_golden import RefRietParams, riet_params
    os.chdir(tempfile.mkdtemp(prefix="tmens_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet.so"))
    rp = RefRietParams(**riet_params(a=a, dt=dt))
    ft = np.array([1e300, 1e300])
    out = []
    for _ in range(nrep):
        rows = np.zeros((256, 7)); nfr = C.c_int()
        n = ref.ref_riet_run(C.byref(rp), C.c_int(L), C.c_int(1), ptr(init), ptr(t_meas), C.c_int(len(t_meas)), ptr(ft), C.c_int(2),
                             None, C.c_uint64(0), None, None, C.c_int(0), C.byref(nfr), ptr(rows), C.c_int(256), None, None)
        out.append(rows[:n, [0, 1, 2, 3, 6]].copy())  # t, <W>, <O>, <P>, active
    return out


def main():
    import oracle_lib as O
    ncpu = os.cpu_count() or 1
    pool = mp.get_context("spawn").Pool(ncpu)
    out = {}
    # DP near criticality (delta_frame_test.cpp constants), rho0 = 1
    L, dt = 64, 0.1
    tm = time_bins(100.0, dt)
    params = (1.0548, 1.0, 0.0, 0.25, 2 ** 0.5, dt, 1.0)
    res = pool.map(dp_job, [(0, L, params, np.ones(L * L), tm, 12)] * ncpu)
    series = np.array([s for r in res for s in r])
    out["ens_dp_plain"] = dict(L=L, params=np.array(params), t_meas=tm, series=series)
    # cubic DP (1storderDP_delta.cpp constants), rho0 = 0.25
    tm = time_bins(40.0, dt)
    params = (-0.72, -2.0, 1.0, 6.25, 1.0, dt, 2.5)
    res = pool.map(dp_job, [(1, L, params, np.full(L * L, 0.25), tm, 8)] * ncpu)
    series = np.array([s for r in res for s in r])
    out["ens_dp_cubic"] = dict(L=L, params=np.array(params), t_meas=tm, series=series)
    # Rietkerk SPB=1 from a fixed mixed frame
    L = 48
    init = O.synthetic_rietkerk_frame(L, seed=17).reshape(3, -1).copy()
    tm = np.concatenate([[0.0], time_bins(30.0, dt)])
    res = pool.map(riet_job, [(L, 0.05, dt, init, tm, 6)] * ncpu)
    series = np.array([s for r in res for s in r])
    out["ens_rietkerk"] = dict(L=L, a=0.05, dt=dt, init=init, t_meas=tm, series=series)
    pool.close()
    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print("wrote", name, d["series"].shape)


if __name__ == "__main__":
    main()
