"""Consumer-side operator fixtures from the UNMODIFIED reference compiled with -DSPB=2 (oracle/_ref/libtm_ref_riet2.so):

  consumer_ops.npz   computeNeighboringSitesCentral (R/:1661), calc_gamma_2Sp_NonRefugia (R/:2530) and
                     RK4_Integrate_Stochastic_2Sp (R/:2472) on seeded inputs with the SPB=2 driver's constants
                     (Rietkerk_FD/order_2stoc_test_rietkerk.cpp:28-94).

Run in the build container only (needs /root/reference):   python tests/golden/make_consumer_golden.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def driver_constants():
    """order_2stoc_test_rietkerk.cpp: vegetation as SPB=1, grazer of 20 kg."""
    dx, dt = 0.1, 0.1
    c, gmax, d, alpha, W0, rW = 10000.0, 0.05 * 10.0 ** -3.0 / 24.0, 0.25 / 24.0, 0.2 / 24.0, 0.2, 0.2 / 24.0
    K1, K2 = 5.0, 5000.0
    D2, D3 = 0.00025 / 24.0, 0.025 / 24.0
    aij = 3.6 * 10.0 ** -6.08 * 20.0 ** -0.37
    hij, ej = 1.0, 0.45
    return dict(dt=dt, dx=dx, a=0.05, c=c, gmax=gmax, alpha=alpha, rW=rW, W0=W0, K1=K1, K2=K2, D2=D2, D3=D3, A01=aij, H01=hij, E1=ej)


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet2.so"))
    assert ref.ref2_spb() == 2
    out = {}
    rng = np.random.default_rng(77)
    # --- disc offsets (sorted by squared distance; ties in libstdc++ std::sort's order) -------------------------------
    for R in (1, 3, 5):
        buf = np.zeros((4 * R * R + 4 * R + 8, 2), dtype=np.int32)
        n = ref.ref2_central_sites(R, buf.ctypes.data_as(C.POINTER(C.c_int)), buf.shape[0])
        out["central_R%d" % R] = buf[:n].copy()
    # --- perception field ------------------------------------------------------------------------------------------------
    cases = {"gamma_full": (40, 5, 1.0, 1.0), "gamma_sparse": (33, 5, 1.0, 0.5), "gamma_small": (24, 3, 0.8, 0.2), "gamma_dead": (16, 3, 1.0, 1.0)}
    for name, (L, R, fr, nveg) in cases.items():
        veg = rng.gamma(2.0, 300.0, L * L); veg[rng.random(L * L) < (1 - nveg) * 0.9] = 0.0
        graz = rng.gamma(2.0, 1.0, L * L)
        g0 = rng.random(L * L)                                  # the array's previous contents (accumulated onto, R/:2581)
        avg = np.array([veg.mean(), graz.mean()])
        if name == "gamma_dead":
            avg[0] = 0.0                                         # no resource left: gamma = 0 (R/:2577)
        g = g0.copy()
        ref.ref2_calc_gamma(C.c_int(L), C.c_int(R), ptr(veg), ptr(graz), ptr(g), ptr(avg), C.c_double(fr), C.c_double(nveg))
        out[name + "_cfg"] = np.array([L, R, fr, nveg])
        for k, a in (("veg", veg), ("g0", g0), ("avg", avg), ("out", g)):
            out[name + "_" + k] = a
    # --- four-field RK4 --------------------------------------------------------------------------------------------------
    k = driver_constants()
    par = np.array([k[x] for x in "dt dx a c gmax alpha rW W0 K1 K2 D2 D3 A01 H01 E1".split()])
    out["rk4_par"] = par
    for name, L, steps in (("rk4_a", 24, 1), ("rk4_b", 17, 5)):
        P = rng.gamma(2.0, 3000.0, L * L); P[rng.random(L * L) < 0.2] = 0.0
        G = rng.gamma(2.0, 2.0, L * L); G[rng.random(L * L) < 0.3] = 0.0
        W = 0.25 * (1 + 0.2 * rng.standard_normal(L * L)); O_ = 6.6 * (1 + 0.2 * rng.standard_normal(L * L))
        f0 = np.ascontiguousarray(np.stack([P, G, W, O_]))
        f = f0.copy()
        for _ in range(steps):
            ref.ref2_rk4_2sp(C.c_int(L), ptr(par), ptr(f))
        out[name + "_cfg"] = np.array([L, steps]); out[name + "_in"] = f0; out[name + "_out"] = f
    np.savez_compressed(os.path.join(HERE, "consumer_ops.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
