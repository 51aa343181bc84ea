"""Golden fixture of the noise-off Rietkerk runner: the UNMODIFIED reference RK4_Wrapper_2D
(Rietkerk_FD/rietkerk_bjork_basic.cpp:1816) compiled into oracle/_ref/libtm_ref_riet.so (oracle/Makefile `ref`; its
random initial-frame routine init_randconstframe is replaced at link time by one that hands it the caller's frames).
Writes tests/golden/riet_det_runner.npz: constants, the r initial frames, t_meas, the runner's own rows."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (only for the synthetic frame rule)

ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet.so"))
dp = C.POINTER(C.c_double)
ref.ref_riet_det_run.argtypes = [dp, C.c_int, C.c_int, dp, dp, C.c_int, dp, C.c_int]

out = {}
for name, L, r, a, dt, t_meas, dead in [("alive", 24, 3, 0.05, 0.1, [1.5, 2.0, 3.0], False),
                                        ("one_dead", 16, 3, 0.05, 0.25, [1.5, 2.5], True)]:
    # SPB=1 driver constants (order_1stoc_test_rietkerk.cpp:14,28-37,48); d = 0.25/24
    par = np.array([a, 1e4, 0.05e-3 / 24, 0.2 / 24, 0.25 / 24, 0.2 / 24, 0.2, 0.00025 / 24, 0.00025 / 24, 0.025 / 24, 5.0, 5000.0,
                    dt, 0.1, 10000.0])
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=300 + j, a=a) for j in range(r)]).reshape(r, 3, L * L)
    if dead:
        frames[1, 0] = 0.0                     # replicate 1 has no vegetation: its rows stop counting as surviving
    tm = np.array(t_meas, dtype=np.float64)
    rows = np.zeros((256, 10))
    n = ref.ref_riet_det_run(par.ctypes.data_as(dp), L, r, np.ascontiguousarray(frames).ctypes.data_as(dp), tm.ctypes.data_as(dp), len(tm),
                             rows.ctypes.data_as(dp), len(rows))
    assert n > 0, n
    out[name + "_par"] = par
    out[name + "_L"] = np.array(L)
    out[name + "_frames"] = frames
    out[name + "_tmeas"] = tm
    out[name + "_rows"] = rows[:n].copy()
    print(name, n, rows[:n, 3], rows[:3])
np.savez_compressed(os.path.join(HERE, "riet_det_runner.npz"), **out)
