"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled by oracle/Makefile (`make ref`).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Every fixture stores the inputs next to the reference's outputs, so the CPU tests can replay the
oracle on them anywhere (the GPU box has no /root/reference).

Replay uniforms come from numpy's legacy RandomState (MT19937 `random_sample`, a frozen stream):
53-bit values k*2^-53, which the replay engine (oracle/ref_shim/replay_urbg.h) hands to the
reference's std:: samplers bit for bit.
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


class RefRietParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in "a c gmax alpha rW W0 d K1 K2 D0 D1 D2 sigma0 dt dx".split()]


def riet_params(a=0.05, dt=0.1):
    D0 = 0.00025 / 24
    return dict(a=a, c=1e4, gmax=0.05e-3 / 24, alpha=0.2 / 24, rW=0.2 / 24, W0=0.2, d=0.25 / 24, K1=5.0, K2=5000.0,
                D0=D0, D1=D0, D2=0.025 / 24, sigma0=(D0 / 0.01) ** 0.5, dt=dt, dx=0.1)


def load_refs():
    d = os.path.join(ROOT, "oracle", "_ref")
    return C.CDLL(os.path.join(d, "libtm_ref_dp.so")), C.CDLL(os.path.join(d, "libtm_ref_riet.so"))


def gen_dp(ref, cubic, L, params, rho0, t_meas, seed, n_u):
    a, b, c, D, sigma, dt, dx = params
    u = uniforms(seed, n_u)
    rows = np.zeros((4096, 6)); cons = C.c_uint64(); ov = C.c_uint64()
    tm = np.asarray(t_meas, dtype=np.float64)
    n = ref.ref_dp_run(C.c_int(int(cubic)), C.c_int(L), *[C.c_double(x) for x in (a, b, c, D, sigma, dt, dx)], C.c_int(1),
                       ptr(rho0), ptr(tm), C.c_int(len(tm)), ptr(u), C.c_uint64(n_u), ptr(rows), C.c_int(4096),
                       C.byref(cons), C.byref(ov))
    assert n > 0 and ov.value == 0, (n, ov.value)
    return dict(cubic=int(cubic), L=L, params=np.array(params), rho0=rho0, t_meas=tm, seed=seed, n_u=n_u,
                rows=rows[:n].copy(), consumed=cons.value)


def gen_riet(ref, L, pd, init, nsteps, seed, n_u):
    rp = RefRietParams(**pd)
    dt = pd["dt"]
    tmeas = np.array([0.0, dt * nsteps])
    frame_t = np.array([dt * k for k in range(1, nsteps + 1)] + [1e300])  # sentinel: the runner reads one past the end (R/:3736)
    u = uniforms(seed, n_u)
    frames = np.zeros((32, 3, L * L)); ftimes = np.zeros(32); nfr = C.c_int(); rows = np.zeros((64, 7))
    cons = C.c_uint64(); ov = C.c_uint64()
    nr = ref.ref_riet_run(C.byref(rp), C.c_int(L), C.c_int(1), ptr(init), ptr(tmeas), C.c_int(len(tmeas)), ptr(frame_t),
                          C.c_int(len(frame_t)), ptr(u), C.c_uint64(n_u), ptr(frames), ptr(ftimes), C.c_int(32), C.byref(nfr),
                          ptr(rows), C.c_int(64), C.byref(cons), C.byref(ov))
    assert nr > 0 and ov.value == 0
    k = nfr.value
    return dict(L=L, params=np.array([pd[n] for n, _ in RefRietParams._fields_]), init=init, nsteps=nsteps, seed=seed, n_u=n_u,
                frames=frames[:k].copy(), frame_times=ftimes[:k].copy(), rows=rows[:nr].copy(), consumed=cons.value)


def main():
    os.chdir(tempfile.mkdtemp(prefix="tmgold_"))
    os.makedirs("../Data/SPDP", exist_ok=True)
    refdp, refr = load_refs()
    import oracle_lib as O
    out = {}

    # --- DP and cubic DP replicate runners, sequential replay ------------------------------------
    L = 16
    rng = np.random.RandomState(5)
    rho = np.ones(L * L); rho[: 4 * L] = 0.0; rho[4 * L: 6 * L] = rng.random_sample(2 * L) * 0.2
    out["dp_plain"] = gen_dp(refdp, False, L, (1.0548, 1.0, 0.0, 0.25, 2 ** 0.5, 0.1, 1.0), rho, [1.5, 2.0, 3.0], 11, 400000)
    rho = np.full(L * L, 0.25); rho[: 3 * L] = 0.0
    out["dp_cubic"] = gen_dp(refdp, True, L, (-0.72, -2.0, 1.0, 6.25, 1.0, 0.1, 2.5), rho, [1.5, 2.0, 3.0], 12, 800000)
    # a sparse, nearly extinct DP state: small Gamma shapes (< 1), Poisson means near 0
    rho = np.zeros(L * L); rho[rng.randint(0, L * L, 12)] = rng.random_sample(12) * 0.05
    out["dp_sparse"] = gen_dp(refdp, False, L, (1.0548, 1.0, 0.0, 0.25, 2 ** 0.5, 0.1, 1.0), rho, [1.2, 1.5], 13, 200000)

    # --- Rietkerk replicate runner: all sampler regimes --------------------------------------------
    L = 16
    init = O.synthetic_rietkerk_frame(L, seed=3).reshape(3, -1).copy()
    out["riet_mixed"] = gen_riet(refr, L, riet_params(), init, 9, 21, 200000)
    init = O.synthetic_rietkerk_frame(L, seed=4, fringe=False).reshape(3, -1).copy()
    out["riet_dense"] = gen_riet(refr, L, riet_params(a=0.07), init, 9, 22, 200000)

    # --- RK4 pieces called directly -------------------------------------------------------------------
    L = 12
    f = O.synthetic_rietkerk_frame(L, seed=6).reshape(3, -1).copy()
    pd = riet_params()
    rp = RefRietParams(**pd)
    g = f.copy(); refr.ref_riet_rk4_stoch(C.byref(rp), C.c_int(L), ptr(g))
    out["rk4_stoch"] = dict(L=L, params=np.array([pd[n] for n, _ in RefRietParams._fields_]), inp=f, out=g)
    g = f.copy(); refr.ref_riet_rk4_det(C.byref(rp), C.c_int(L), C.c_int(25), ptr(g))
    out["rk4_det"] = dict(L=L, params=np.array([pd[n] for n, _ in RefRietParams._fields_]), inp=f, out=g, nsteps=25)
    rho = np.random.RandomState(8).random_sample(L * L) * 1.5
    g = rho.copy(); refdp.ref_dp_rk4(C.c_int(L), *[C.c_double(x) for x in (-0.72, -2.0, 1.0, 6.25, 0.0, 0.1, 2.5)], ptr(g))
    out["dp_rk4"] = dict(L=L, b=-2.0, c=1.0, dt=0.1, inp=rho, out=g)

    # --- bookkeeping: time bins, reductions -------------------------------------------------------------
    bins = {}
    for t_max, dt in ((100.0, 0.1), (2511.9, 0.1), (25.0, 0.25), (1000.0, 0.12)):
        buf = np.zeros(4096)
        n = refdp.ref_dp_time_bins(C.c_double(t_max), C.c_double(dt), ptr(buf), C.c_int(4096))
        bins["%g_%g" % (t_max, dt)] = buf[:n].copy()
    out["time_bins"] = bins
    v = np.random.RandomState(9).random_sample(1000) - 0.3
    refr.ref_riet_mean.restype = C.c_double; refr.ref_riet_occupied.restype = C.c_double
    out["reductions"] = dict(v=v, mean=refr.ref_riet_mean(ptr(v), C.c_int(v.size)), occupied=refr.ref_riet_occupied(ptr(v), C.c_int(v.size)))

    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print("wrote", name, {k: (np.asarray(v).shape if hasattr(v, "shape") else v) for k, v in d.items() if k not in ("rho0", "init")})


if __name__ == "__main__":
    main()
