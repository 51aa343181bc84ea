"""FKE movement fixtures from the UNMODIFIED reference (oracle/_ref/libtm_ref_fke.so, `make -C oracle ref`):

  fke_move.npz   set_global_FKE_params (F/:195), precompute_gaussian (F/:1735) and advdiff_FKE_MultiSp (F/:3614)
                 of Rietkerk_FPE compiled with SPB=2: derived movement constants of the SPB=2 driver
                 (order_2stoc_unity_rietkerk.cpp:50-69,202), Gaussian stencils, and the moved grazer field for
                 seeded random inputs in four window/displacement regimes.

Run in the build container only (needs /root/reference):   python tests/golden/make_fke_golden.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def inputs(L, seed, dead_frac, vmax):
    rng = np.random.default_rng(seed)
    rho = rng.gamma(2.0, 3.0, L * L)
    rho[rng.random(L * L) < dead_frac] = 0.0
    rho[rng.random(L * L) < 0.02] = 5e-11            # below epsilon = 1e-10: skipped, F/:3622
    gamma = rng.random(L * L)
    gamma[rng.random(L * L) < 0.05] = 0.0
    gamma[rng.random(L * L) < 0.05] = 1.0
    v = (rng.random((L * L, 2)) * 2 - 1) * vmax
    return rho, gamma, np.ascontiguousarray(v)


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke.so"))
    assert ref.ref_fke_spb() == 2
    out = {}
    # --- the SPB=2 driver's constants: order_2stoc_unity_rietkerk.cpp:50-69 ---------------------------------
    dx, d0, d1, d2, d3, v1 = 0.1, 0.00025 / 24.0, 0.01028, 0.00025 / 24.0, 0.025 / 24.0, 0.45427
    dtv1 = round(0.11817455 * 100.0) / 100.0
    D = np.array([d0, d1, d2, d3]); v = np.array([0.0, v1]); dtadv = np.array([dtv1, dtv1])
    L = 48
    derived = np.zeros(8)
    ref.ref_fke_set_params(ptr(D), ptr(v), ptr(dtadv), C.c_double(dtv1), C.c_double(dx), C.c_int(L), C.c_int(4), ptr(derived))
    out["driver_inputs"] = np.array([d1, v1, dtv1, dx, 4.0])
    out["driver_derived"] = derived
    bD, bV1, bV2 = int(derived[4]), int(derived[5]), int(derived[6])
    # stencils as the runner builds them, F/:3769-3771 (sigma[s] = 1 for the grazer, driver line 54)
    for name, (sig, R) in {"g_s1_R1": (1.0, 1), "g_s1_R3": (1.0, 3), "g_s07_R5": (0.7, 5), "g_s25_R9": (2.5, 9)}.items():
        st = np.zeros(R + 1)
        ref.ref_fke_precompute_gaussian(C.c_double(sig), C.c_int(R), ptr(st))
        out[name] = st
    cases = {
        # name: (L, bound_D, v_first, v_second, prefactor, sigma_D_stencil, sigma_V_stencil, vmax, dead_frac, seed)
        "driver": (L, bD, bV1, bV2, derived[3], 1.0, 1.0, 0.9, 0.2, 11),
        "wide": (40, 3, 3, 2, 2.5, 1.0, 0.7, 1.3, 0.3, 12),           # asymmetric (lo, hi) = (first, second), F/:3640
        "wrap": (9, 4, 4, 4, 3.0, 2.5, 2.5, 1.4, 0.1, 13),            # windows of 9 = L: every target hit once per term
        "sparse": (64, 2, 1, 1, 1.0, 1.0, 1.0, 2.4, 0.97, 14),         # mostly dead sources
    }
    for name, (L, bD, f1, f2, pref, sD, sV, vmax, dead, seed) in cases.items():
        ref.ref_fke_override(C.c_int(bD), C.c_int(f1), C.c_int(f2), C.c_double(pref), C.c_int(L))
        gD = np.zeros(bD + 1); gV = np.zeros(max(f1, f2) + 1)
        ref.ref_fke_precompute_gaussian(C.c_double(sD), C.c_int(bD), ptr(gD))
        ref.ref_fke_precompute_gaussian(C.c_double(sV), C.c_int(max(f1, f2)), ptr(gV))
        rho, gamma, veff = inputs(L, seed, dead, vmax)
        moved = np.zeros(L * L)
        ref.ref_fke_advdiff(C.c_int(L), ptr(rho), ptr(gamma), ptr(veff), ptr(gD), C.c_int(gD.size), ptr(gV), C.c_int(gV.size), ptr(moved))
        out[name + "_cfg"] = np.array([L, bD, f1, f2, pref, sD, sV, 1e-10])
        for k, a in (("rho", rho), ("gamma", gamma), ("veff", veff), ("gD", gD), ("gV", gV), ("out", moved)):
            out[name + "_" + k] = a
        print(name, "mass in", rho[rho >= 1e-10].sum(), "mass out", moved.sum())
    np.savez_compressed(os.path.join(HERE, "fke_move.npz"), **out)


if __name__ == "__main__":
    main()
