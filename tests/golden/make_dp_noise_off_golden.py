"""Golden fixture of the noise-off DP runner: the UNMODIFIED reference percNoDiff_2D (Percolation_FD/stochasticSPDP.cpp:549)
compiled into oracle/_ref/libtm_ref_dp.so (oracle/Makefile `ref`), run here where /root/reference is mounted.
Writes tests/golden/dp_noise_off.npz: parameters, initial field, the runner's own rows {a, replicate, t, cell, rho}."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_dp.so"))
dp = C.POINTER(C.c_double)
ref.ref_dp_noise_off.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, dp, C.c_int, C.c_double, dp, C.c_int]

out = {}
for name, L, a, b, D, dt, dx, seed, t_stop in [
        ("growth", 24, 1.2, 1.0, 0.25, 0.1, 1.0, 1, [0.5, 1.0, 2.0]),
        ("decay_patchy", 20, -0.72, -2.0, 6.25, 0.01, 2.5, 2, [0.05, 0.1, 0.3]),
]:
    rng = np.random.default_rng(seed)
    rho0 = rng.random(L * L)
    rho0[rng.random(L * L) < 0.4] = 0.0                      # empty patches: the Laplacian fills them
    ts = np.array(t_stop, dtype=np.float64)
    rows = np.zeros((len(ts) * L * L + 16, 5))
    n = ref.ref_dp_noise_off(L, a, b, D, dt, dx, rho0.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts), float(ts[-1]),
                             rows.ctypes.data_as(dp), len(rows))
    assert n == len(ts) * L * L, n
    out[name + "_par"] = np.array([L, a, b, D, dt, dx])
    out[name + "_rho0"] = rho0
    out[name + "_tstop"] = ts
    out[name + "_rows"] = rows[:n].copy()
np.savez_compressed(os.path.join(HERE, "dp_noise_off.npz"), **out)
print("wrote dp_noise_off.npz", {k: v.shape for k, v in out.items()})
