"""SPB = 3 (vegetation + grazer + predator) fixtures from the UNMODIFIED reference (oracle/_ref/libtm_ref_riet3.so:
Rietkerk_FD/rietkerk_bjork_basic.cpp compiled with -DSPB=3):

  runner_3sp.npz   rietkerk_Dornic_2D_MultiSp (R/:3397), one replicate under sequential replay of a seeded uniform sequence: all five
                   fields (P, G, Pr, W, O) at full precision at t = 0 and after the first steps, the uniforms consumed;
                   calc_gamma_3Sp_NonRefugia (R/:3153) and RK4_Integrate_Stochastic_MultiSp (R/:2652) on their own; the disc list.
                   Constants: order_3stoc_test_rietkerk.cpp:64-111 (the 20 kg grazer / 100 kg predator block).

Run in the build container only (needs /root/reference):   python tests/golden/make_runner3_golden.py
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
from make_runner2_golden import ptr, uniforms   # noqa: E402


def driver_par(a=0.05, dt=0.1, dx=0.1):
    mG, mP = 20.0, 100.0
    d0, d1, d2, d3, d4 = 0.00025 / 24.0, 0.0298, 0.05221, 0.00025 / 24.0, 0.025 / 24.0
    aij = 3.6 * 10.0 ** -6.08 * mG ** -0.37
    ajm = 3.6 * 10.0 ** -6.08 * mP ** -0.37
    m1 = 0.061609 * mG ** -0.25 / 8760.0
    m2 = 0.061609 * mP ** -0.25 / 8760.0
    # {a, c, gmax, alpha, rW, W0, d, K1, K2,  D0..D4,  sigma0, sigma1, sigma2,  v1, v2,  aij, hij, ej,  ajm, hjm, em,  m1, m2,  pR1, pR2,  dt, dx}
    return np.array([a, 10000.0, 0.05 * 10.0 ** -3.0 / 24.0, 0.2 / 24.0, 0.2 / 24.0, 0.2, 0.25 / 24.0, 5.0, 5000.0,
                     d0, d1, d2, d3, d4, (d0 / (dx * dx)) ** 0.5, 1.0, 3.0, 0.45427, 0.40587, aij, 1.0, 0.45, ajm, 1.0, 0.85,
                     m1, m2, 1.04285 / 2.0, 1.25536 / 2.0, dt, dx])


def frame(L, seed, predators=True, grazers=True):
    import oracle_lib as O
    rng = np.random.default_rng(seed)
    f3 = O.synthetic_rietkerk_frame(L, seed=seed).reshape(3, -1)
    G = rng.gamma(2.0, 1.5, L * L)
    G[rng.random(L * L) < 0.3] = 0.0
    G.reshape(L, L)[:, : L // 3] = 0.0
    G.reshape(L, L)[:, L // 3: L // 3 + 2] = 10 ** rng.uniform(-5, -1, (L, 2))
    R = rng.gamma(1.5, 0.4, L * L)
    R[rng.random(L * L) < 0.4] = 0.0
    R.reshape(L, L)[: L // 4, :] = 0.0
    if not grazers:
        G[:] = 0.0
    if not predators:
        R[:] = 0.0
    return np.ascontiguousarray(np.stack([f3[0], G, R, f3[1], f3[2]]))


def main():
    os.chdir(tempfile.mkdtemp(prefix="tmgold3_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet3.so"))
    assert ref.ref3_spb() == 3
    out = {}
    for name, L, nsteps, seed, pred, graz in (("mixed", 20, 5, 51, True, True), ("no_predators", 16, 3, 52, False, True),
                                              ("no_grazers", 16, 3, 53, True, False)):
        par = driver_par()
        dt = par[29]
        init = frame(L, seed, pred, graz)
        n_u = 600000
        u = uniforms(seed, n_u)
        tmeas = np.array([0.0, dt * nsteps])
        frame_t = np.array([dt * k for k in range(1, nsteps + 1)] + [1e300])
        frames = np.zeros((32, 5, L * L)); ftimes = np.zeros(32); cons = C.c_uint64(); ov = C.c_uint64()
        k = ref.ref3_run(ptr(par), C.c_int(L), ptr(init), ptr(tmeas), C.c_int(len(tmeas)), ptr(frame_t), C.c_int(len(frame_t)),
                         ptr(u), C.c_uint64(n_u), ptr(frames), ptr(ftimes), C.c_int(32), C.byref(cons), C.byref(ov))
        assert k > 1 and ov.value == 0, (k, ov.value)
        print(name, "frames", k, "times", ftimes[:k], "consumed", cons.value)
        out[name + "_par"] = par; out[name + "_L"] = np.array(L); out[name + "_seed"] = np.array([seed, n_u])
        out[name + "_init"] = init; out[name + "_frames"] = frames[:k].copy(); out[name + "_times"] = ftimes[:k].copy()
        out[name + "_consumed"] = np.array(cons.value)
    par = driver_par()
    R = int(max(par[27], par[28]) / par[30])
    cen = np.zeros((4096, 2), dtype=np.int32)
    n = ref.ref3_central_sites(C.c_int(R), cen.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(4096))
    out["central_R%d" % R] = cen[:n].copy()
    # the two operators on their own
    L = 18
    f = frame(L, 61)
    fr = np.array([par[27] / (par[30] * (max(par[27], par[28]) / par[30])), par[28] / (par[30] * (max(par[27], par[28]) / par[30]))])
    for tag, avg in (("all", [f[0].mean(), f[1].mean(), f[2].mean()]), ("nopred", [f[0].mean(), f[1].mean(), 0.0])):
        gam = np.random.default_rng(5).random((2, L * L))          # entry values must not matter (assigned, not accumulated)
        avg = np.array(avg)
        ref.ref3_calc_gamma(C.c_int(L), C.c_int(R), ptr(np.ascontiguousarray(f[:3])), ptr(gam), ptr(avg), ptr(fr), C.c_double(0.62))
        out["gamma_" + tag] = gam.copy(); out["gamma_" + tag + "_avg"] = avg
    out["gamma_fields"] = f[:3].copy(); out["gamma_fr"] = fr; out["gamma_nveg"] = np.array(0.62); out["gamma_L"] = np.array(L)
    rk = np.array([par[29], par[30], par[0], par[1], par[2], par[3], par[4], par[5], par[7], par[8], par[12], par[13],
                   par[19], par[20], par[21], par[22], par[23], par[24]])
    g = f.copy()
    for _ in range(3):
        ref.ref3_rk4(C.c_int(L), ptr(rk), ptr(g))
    out["rk4_par"] = rk; out["rk4_in"] = f.copy(); out["rk4_out3"] = g.copy()
    np.savez_compressed(os.path.join(HERE, "runner_3sp.npz"), **out)


if __name__ == "__main__":
    main()
