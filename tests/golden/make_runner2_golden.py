"""SPB = 2 (vegetation + grazer) replicate-runner fixtures from the UNMODIFIED reference (oracle/_ref/libtm_ref_riet2.so):

  runner_2sp.npz   rietkerk_Dornic_2D_MultiSp (R/:3397) compiled with -DSPB=2, one replicate under sequential replay of a
                   seeded uniform sequence: all four fields at full precision at t = 0 and after every one of the first
                   steps, and the number of uniforms the runner consumed.  Constants: order_2stoc_test_rietkerk.cpp:28-94.

Run in the build container only (needs /root/reference):   python tests/golden/make_runner2_golden.py
"""
import ctypes as C
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


def uniforms(seed, n):
    return np.random.RandomState(seed).random_sample(n)


def driver_par(a=0.05, dt=0.1, dx=0.1):
    d0, d1, d2, d3 = 0.00025 / 24.0, 0.01028, 0.00025 / 24.0, 0.025 / 24.0
    aij = 3.6 * 10.0 ** -6.08 * 20.0 ** -0.37
    m1 = 0.061609 * 20.0 ** -0.25 / 8760.0
    # {a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, sigma0, sigma1, v1, aij, hij, ej, m1, pR1, dt, dx}
    return np.array([a, 10000.0, 0.05 * 10.0 ** -3.0 / 24.0, 0.2 / 24.0, 0.2 / 24.0, 0.2, 0.25 / 24.0, 5.0, 5000.0, d0, d1, d2, d3,
                     (d0 / (dx * dx)) ** 0.5, 1.0, 0.45427, aij, 1.0, 0.45, m1, 1.04285 / 2.0, dt, dx])


def frame(L, seed, grazers=True):
    import oracle_lib as O
    rng = np.random.default_rng(seed)
    f3 = O.synthetic_rietkerk_frame(L, seed=seed).reshape(3, -1)
    G = rng.gamma(2.0, 1.5, L * L)
    G[rng.random(L * L) < 0.3] = 0.0
    G.reshape(L, L)[:, : L // 3] = 0.0                         # an empty region: the skip rule of R/:4048
    G.reshape(L, L)[:, L // 3: L // 3 + 2] = 10 ** rng.uniform(-5, -1, (L, 2))
    if not grazers:
        G[:] = 0.0
    return np.ascontiguousarray(np.stack([f3[0], G, f3[1], f3[2]]))


def main():
    os.chdir(tempfile.mkdtemp(prefix="tmgold2_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet2.so"))
    assert ref.ref2_spb() == 2
    out = {}
    for name, L, nsteps, seed, grazers in (("mixed", 20, 6, 31, True), ("no_grazers", 16, 4, 32, False)):
        par = driver_par()
        dt = par[21]
        init = frame(L, seed, grazers)
        n_u = 400000
        u = uniforms(seed, n_u)
        tmeas = np.array([0.0, dt * nsteps])
        frame_t = np.array([dt * k for k in range(1, nsteps + 1)] + [1e300])   # sentinel: the runner reads one past the end (R/:3736)
        frames = np.zeros((32, 4, L * L)); ftimes = np.zeros(32); cons = C.c_uint64(); ov = C.c_uint64()
        k = ref.ref2_run(ptr(par), C.c_int(L), ptr(init), ptr(tmeas), C.c_int(len(tmeas)), ptr(frame_t), C.c_int(len(frame_t)),
                         ptr(u), C.c_uint64(n_u), ptr(frames), ptr(ftimes), C.c_int(32), C.byref(cons), C.byref(ov))
        assert k > 1 and ov.value == 0, (k, ov.value)
        print(name, "frames", k, "times", ftimes[:k], "consumed", cons.value)
        out[name + "_par"] = par; out[name + "_L"] = np.array(L); out[name + "_seed"] = np.array([seed, n_u])
        out[name + "_init"] = init; out[name + "_frames"] = frames[:k].copy(); out[name + "_times"] = ftimes[:k].copy()
        out[name + "_consumed"] = np.array(cons.value)
    np.savez_compressed(os.path.join(HERE, "runner_2sp.npz"), **out)


if __name__ == "__main__":
    main()
