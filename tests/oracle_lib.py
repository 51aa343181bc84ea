"""ctypes view of the CPU oracle (oracle/tm_oracle.h).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
dp = C.POINTER(C.c_double)

SRC_SEQ, SRC_REPLAY, SRC_PHILOX = 0, 1, 2


class OParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("L", C.c_int32)] + [
        (n, C.c_double) for n in (
            "dt dx a b c D sigma rc gmax alpha rW W0 d K1 K2 D0 D1 D2 sigma0 clamp_eps clamp_lo").split()]


class ORng(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("u", dp), ("n_u", C.c_uint64), ("pos", C.c_uint64),
                ("stride", C.c_uint64), ("seed", C.c_uint64), ("trial", C.c_uint32), ("step", C.c_uint32),
                ("overflow", C.c_uint64), ("draws", C.c_uint64), ("norm_saved_available", C.c_int32),
                ("_pad2", C.c_int32), ("norm_saved", C.c_double)]


class OFkeTerms(C.Structure):
    """struct tmo_fke_terms."""
    _fields_ = [("L", C.c_int32), ("bound_D", C.c_int32), ("v_row_lo", C.c_int32), ("v_row_hi", C.c_int32),
                ("v_col_lo", C.c_int32), ("v_col_hi", C.c_int32), ("mu_vel_prefactor", C.c_double), ("epsilon", C.c_double)]


class OFkeDerived(C.Structure):
    """struct tmo_fke_derived."""
    _fields_ = [("beta_vel", C.c_double), ("sigma_D", C.c_double), ("sigma_vD", C.c_double), ("mu_vel_prefactor", C.c_double),
                ("bound_D", C.c_int32), ("bound_vD_first", C.c_int32), ("bound_vD_second", C.c_int32), ("pad_", C.c_int32)]


class OConsParams(C.Structure):
    """struct tmo_cons_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in "dt a c gmax alpha rW W0 K1 K2 Dw Do A01 H01 E1".split()]


class OCons3Params(C.Structure):
    """struct tmo_cons3_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in
                                                         "dt a c gmax alpha rW W0 K1 K2 Dw Do A01 H01 E1 A12 H12 E2".split()]


class OGrazParams(C.Structure):
    """struct tmo_graz_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in "dt dx D v M sigma D0floor".split()]


_lib = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def oracle():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libtm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        for n in ("tmo_log", "tmo_exp", "tmo_lgamma_int"):
            getattr(L, n).restype = C.c_double
            getattr(L, n).argtypes = [C.c_double]
        L.tmo_mean.restype = C.c_double
        L.tmo_mean.argtypes = [dp, C.c_int64]
        L.tmo_occupied.restype = C.c_double
        L.tmo_occupied.argtypes = [dp, C.c_int64]
        L.tmo_step.argtypes = [C.POINTER(OParams), C.POINTER(ORng), dp]
        L.tmo_rk4.argtypes = [C.POINTER(OParams), dp, C.c_int]
        L.tmo_dornic_only.argtypes = [C.POINTER(OParams), C.POINTER(ORng), dp, dp]
        L.tmo_u52_from_bits.restype = C.c_double
        L.tmo_u52_from_bits.argtypes = [C.c_uint64]
        L.tmo_test_normal_pair.restype = C.c_double
        L.tmo_test_normal_pair.argtypes = [dp, C.c_uint64, C.POINTER(C.c_uint64), dp]
        L.tmo_test_normal_pair_philox.restype = C.c_double
        L.tmo_test_normal_pair_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.tmo_sincos_turn128.restype = None
        L.tmo_sincos_turn128.argtypes = [C.c_double, dp, dp]
        L.tmo_test_poisson_philox.restype = C.c_long
        L.tmo_test_poisson_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        L.tmo_test_poisson.restype = C.c_long
        L.tmo_test_poisson.argtypes = [dp, C.c_uint64, C.POINTER(C.c_uint64), C.c_double]
        L.tmo_test_gamma.restype = C.c_double
        L.tmo_test_gamma.argtypes = [dp, C.c_uint64, C.POINTER(C.c_uint64), C.c_double]
        L.tmo_logarithm10_time_bins.argtypes = [C.c_double, C.c_double, dp, C.c_int]
        L.tmo_spread_stats.restype = None
        L.tmo_spread_stats.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.tmo_spreading_incremental_surv_runs.restype = None
        L.tmo_spreading_incremental_surv_runs.argtypes = [dp, dp, C.c_int]
        L.tmo_var_mean_incremental_surv_runs.argtypes = [dp, dp, C.c_int, C.c_int]
        L.tmo_fke_precompute_gaussian.restype = None
        L.tmo_fke_precompute_gaussian.argtypes = [C.c_double, C.c_int, dp]
        L.tmo_fke_params.restype = None
        L.tmo_fke_params.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(OFkeDerived)]
        L.tmo_fke_move.argtypes = [C.POINTER(OFkeTerms), dp, dp, dp, dp, dp, dp]
        L.tmo_gamma_perception.restype = None
        L.tmo_gamma_perception.argtypes = [C.c_int, dp, C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double, dp]
        L.tmo_rk4_2sp.argtypes = [C.POINTER(OConsParams), dp]
        L.tmo_grazer_dornic.argtypes = [C.POINTER(OGrazParams), C.POINTER(ORng), dp, dp, dp, dp, C.c_int]
        L.tmo_step_2sp.argtypes = [C.POINTER(OParams), C.POINTER(OGrazParams), C.POINTER(OConsParams), C.POINTER(C.c_int32), C.c_int,
                                   C.c_double, C.POINTER(ORng), dp, dp, dp, C.c_int]
        L.tmo_step_2sp_means.argtypes = L.tmo_step_2sp.argtypes + [dp]
        L.tmo_gamma_3sp.argtypes = [C.c_int, dp, dp, dp, C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double, dp, dp, dp]
        L.tmo_gamma_3sp.restype = None
        L.tmo_rk4_3sp.argtypes = [C.POINTER(OCons3Params), dp]
        L.tmo_step_3sp.argtypes = [C.POINTER(OParams), C.POINTER(OGrazParams), C.POINTER(OGrazParams), C.POINTER(OCons3Params),
                                   C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double, C.POINTER(ORng), dp, dp, dp, C.c_int, dp]
        L.tmo_rk4_fpe_3sp.argtypes = [C.POINTER(OCons3Params), C.c_double, C.c_double, dp]
        L.tmo_step_fpe_3sp.argtypes = [C.POINTER(OParams), C.POINTER(OCons3Params), dp, dp, C.POINTER(OFkeTerms), C.POINTER(dp), C.POINTER(dp),
                                       C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double, C.POINTER(ORng), dp, dp, dp, C.c_int, dp]
        L.tmo_step_fpe_2sp.argtypes = [C.POINTER(OParams), C.POINTER(OConsParams), C.c_double, C.POINTER(OFkeTerms), dp, dp,
                                       C.POINTER(C.c_int32), C.c_int, C.c_double, C.POINTER(ORng), dp, dp, dp, dp]
        L.tmo_rk4_fpe_2sp.argtypes = [C.POINTER(OConsParams), C.c_double, dp]
        L.tmo_gamma_perception_fpe.argtypes = [C.c_int, dp, C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double, dp]
        L.tmo_gamma_perception_fpe.restype = None
        L.tmo_dp_noise_off_step.argtypes = [C.c_int] + [C.c_double] * 5 + [dp]
        _lib = L
    return _lib


def to_oparams(p):
    """Product Params -> oracle params (same field names; separate structs on purpose)."""
    o = OParams()
    for name, _ in OParams._fields_:
        setattr(o, name, getattr(p, name))
    return o


def ptr(a):
    return a.ctypes.data_as(dp)


def make_rng(kind, u=None, stride=0, seed=0, trial=0, step=0, pos=0):
    g = ORng()
    g.kind = kind
    if u is not None:
        g.u = ptr(u)
        g.n_u = u.size
    g.pos, g.stride, g.seed, g.trial, g.step = pos, stride, seed, trial, step
    return g


def step(p, rng, fields):
    """One oracle step in place on fields [n_fields][L*L] (contiguous float64)."""
    assert fields.flags["C_CONTIGUOUS"] and fields.dtype == np.float64
    rc = oracle().tmo_step(C.byref(p), C.byref(rng) if rng is not None else None, ptr(fields))
    assert rc == 0
    return fields


def dornic_only(p, rng, prev):
    out = np.empty_like(prev)
    rc = oracle().tmo_dornic_only(C.byref(p), C.byref(rng), ptr(prev), ptr(out))
    assert rc == 0
    return out


def mean(x):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    return oracle().tmo_mean(ptr(x), x.size)


def occupied(x):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    return oracle().tmo_occupied(ptr(x), x.size)


def spread_stats(x, g, c_x, c_y):
    """(sum of squared distances, N, border hit) of one g x g field about (c_x, c_y): P/:156 and P/:380."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    sq, n, b = C.c_double(), C.c_int64(), C.c_int()
    oracle().tmo_spread_stats(ptr(x), g, c_x, c_y, C.byref(sq), C.byref(n), C.byref(b))
    return sq.value, n.value, b.value


def synthetic_rietkerk_frame(L, seed=1234, a=0.05, dP=10000.0, fringe=True):
    """MFT speckle frame of SURVEY 8(d) (init_exprtk_randbiMFTframe rule, R/:474-496), plus an
    optional empty/fringe band so that every sampler regime is exercised."""
    rng = np.random.default_rng(seed)
    n = L * L
    c, d, rW, Wstar = 1e4, 0.25 / 24, 0.2 / 24, 5.0
    Vstar = (c / d) * a - rW * Wstar * c / d
    P = np.full(n, Vstar)
    P[rng.random(n) < 0.015] = max(dP, 5 * Vstar)
    P = P.reshape(L, L)
    if fringe:
        w = max(2, L // 8)
        P[:, :w] = 0.0
        P[:, w:w + max(2, L // 16)] = 10 ** rng.uniform(-5, -1, (L, max(2, L // 16)))
    f = np.empty((3, L, L))
    f[0] = P
    f[1] = Wstar * a
    f[2] = 120.4744181 * a + 0.58243061
    return f


def fke_gaussian(sigma, R):
    """precompute_gaussian, F/:1735."""
    out = np.zeros(R + 1)
    oracle().tmo_fke_precompute_gaussian(sigma, R, ptr(out))
    return out


def fke_params(D, v, Dt, dx, CI=4):
    """set_global_FKE_params for one species, F/:195-262."""
    d = OFkeDerived()
    oracle().tmo_fke_params(D, v, Dt, dx, CI, C.byref(d))
    return d


def fke_move(L, bound_D, rows, cols, prefactor, eps, rho, gamma, veff, gD, gV):
    """advdiff_FKE_MultiSp for one species, F/:3614; rows/cols = (lo, hi) of the advection window."""
    t = OFkeTerms(L, bound_D, rows[0], rows[1], cols[0], cols[1], prefactor, eps)
    rho, gamma, veff = (np.ascontiguousarray(a, dtype=np.float64) for a in (rho, gamma, veff))
    gD, gV = np.ascontiguousarray(gD, dtype=np.float64), np.ascontiguousarray(gV, dtype=np.float64)
    assert gD.size >= bound_D + 1 and gV.size >= max(rows + cols) + 1
    out = np.empty(L * L)
    rc = oracle().tmo_fke_move(C.byref(t), ptr(rho), ptr(gamma), ptr(veff), ptr(gD), ptr(gV), ptr(out))
    assert rc == 0
    return out


def cons_params(L, par):
    """par = {dt, dx, a, c, gmax, alpha, rW, W0, K1, K2, D2, D3, A01, H01, E1} (tests/golden/make_consumer_golden.py)."""
    dt, dx, a, c, gmax, alpha, rW, W0, K1, K2, D2, D3, A01, H01, E1 = [float(x) for x in par]
    return OConsParams(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D2 / (dx * dx), D3 / (dx * dx), A01, H01, E1)


def gamma_window(n_central, fr, nveg_frac):
    """(new_length, eff) of calc_gamma_2Sp_NonRefugia, R/:2533-2540, :2569."""
    nveg_frac = max(nveg_frac, 0.35)
    new_length = max(int(nveg_frac * nveg_frac * n_central), 1)
    return new_length, int(fr * fr * new_length)


def gamma_perception(L, prey, offsets, eff, avg_prey, avg_self, gamma0):
    g = np.ascontiguousarray(gamma0, dtype=np.float64).copy()
    prey = np.ascontiguousarray(prey, dtype=np.float64)
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    oracle().tmo_gamma_perception(L, ptr(prey), off.ctypes.data_as(C.POINTER(C.c_int32)), eff, avg_prey, avg_self, ptr(g))
    return g


def rk4_2sp(p, fields, steps=1):
    f = np.ascontiguousarray(fields, dtype=np.float64).copy()
    for _ in range(steps):
        assert oracle().tmo_rk4_2sp(C.byref(p), ptr(f)) == 0
    return f


def grazer_dornic(p, rng, prev, gamma, trig_libm=False):
    """-> (rho*, gamma clamped, v_eff [L*L, 2]); the grazer's Dornic update, R/:3908-4114."""
    n = p.L * p.L
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    g = np.ascontiguousarray(gamma, dtype=np.float64).copy()
    out, veff = np.empty(n), np.empty((n, 2))
    assert oracle().tmo_grazer_dornic(C.byref(p), C.byref(rng), ptr(prev), ptr(g), ptr(out), ptr(veff), 1 if trig_libm else 0) == 0
    return out, g, veff


def runner2_params(L, par):
    """(vegetation, grazer, RK4) oracle parameter blocks from the SPB = 2 runner vector of tests/golden/make_runner2_golden.py:
    {a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, sigma0, sigma1, v1, aij, hij, ej, m1, pR1, dt, dx}."""
    a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, s0, s1, v1, aij, hij, ej, m1, pR1, dt, dx = [float(x) for x in par]
    pv = OParams()
    pv.model, pv.L, pv.dt, pv.dx, pv.a = 2, L, dt, dx, a
    pv.rc, pv.gmax, pv.alpha, pv.rW, pv.W0, pv.d, pv.K1, pv.K2 = c, gmax, alpha, rW, W0, d, K1, K2
    pv.D0, pv.D1, pv.D2, pv.sigma0, pv.clamp_eps, pv.clamp_lo = D0, D2, D3, s0, 1e-12, -1e-80
    pg = OGrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    pc = OConsParams(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D2 / (dx * dx), D3 / (dx * dx), aij, hij, ej)
    r_max = pR1 / dx                       # r_max_effective = max(pR)/dx, R/:3405
    return pv, pg, pc, int(r_max), pR1 / (dx * r_max)


def step_2sp(pv, pg, pc, central, fr, rng, f, gamma, trig_libm=False, means=None):
    """One SPB = 2 runner step in place on f [4][L*L] and gamma [L*L]; returns the grazer's advection vectors.
    means = (<P>, <G>) of the entry state replaces the left-to-right sums (tmo_step_2sp_means)."""
    cen = np.ascontiguousarray(central, dtype=np.int32)
    veff = np.zeros((pv.L * pv.L, 2))
    m = None if means is None else np.array([means[0], means[1]], dtype=np.float64)
    rc = oracle().tmo_step_2sp_means(C.byref(pv), C.byref(pg), C.byref(pc), cen.ctypes.data_as(C.POINTER(C.c_int32)), len(cen), fr,
                                     C.byref(rng), ptr(f), ptr(gamma), ptr(veff), 1 if trig_libm else 0, None if m is None else ptr(m))
    assert rc == 0
    return veff


def runner3_params(L, par):
    """(vegetation, grazer, predator, RK4) oracle parameter blocks from the SPB = 3 runner vector of tests/golden/make_runner3_golden.py:
    {a, c, gmax, alpha, rW, W0, d, K1, K2, D0..D4, sigma0..2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2, dt, dx}.
    Returns (pv, pg1, pg2, pc, disc radius, fr1, fr2)."""
    (a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, D4, s0, s1, s2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2,
     dt, dx) = [float(x) for x in par]
    pv = OParams()
    pv.model, pv.L, pv.dt, pv.dx, pv.a = 2, L, dt, dx, a
    pv.rc, pv.gmax, pv.alpha, pv.rW, pv.W0, pv.d, pv.K1, pv.K2 = c, gmax, alpha, rW, W0, d, K1, K2
    pv.D0, pv.D1, pv.D2, pv.sigma0, pv.clamp_eps, pv.clamp_lo = D0, D3, D4, s0, 1e-12, -1e-80
    pg1 = OGrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    pg2 = OGrazParams(L, 0, dt, dx, D2, v2, m2, s2, D0 / (dx * dx))
    pc = OCons3Params(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), aij, hij, ej, ajm, hjm, em)
    r_max = max(pR1, pR2) / dx                 # r_max_effective, R/:3405
    return pv, pg1, pg2, pc, int(r_max), pR1 / (dx * r_max), pR2 / (dx * r_max)


def step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, rng, f, gamma, trig_libm=False, means=None):
    """One SPB = 3 runner step in place on f [5][L*L] and gamma [2][L*L]; returns the two consumers' advection vectors [2][L*L][2]."""
    cen = np.ascontiguousarray(central, dtype=np.int32)
    veff = np.zeros((2, pv.L * pv.L, 2))
    m = None if means is None else np.array(means, dtype=np.float64)
    rc = oracle().tmo_step_3sp(C.byref(pv), C.byref(pg1), C.byref(pg2), C.byref(pc), cen.ctypes.data_as(C.POINTER(C.c_int32)), len(cen),
                               fr1, fr2, C.byref(rng), ptr(f), ptr(gamma), ptr(veff), 1 if trig_libm else 0, None if m is None else ptr(m))
    assert rc == 0
    return veff


def fpe_runner3_params(L, par, derived):
    """Oracle parameter blocks of the Rietkerk_FPE SPB = 3 runner from the vector of tests/golden/make_fpe_runner3_golden.py (the SPB = 3
    runner vector plus dtAdv1, dtAdv2) and the [2][8] constants set_global_FKE_params derived for grazer and predator.
    Returns (pv, pc, M[2], v[2], terms (OFkeTerms * 2), gD[2], gV[2], disc radius, fr1, fr2)."""
    pv, pg1, pg2, pc, R, fr1, fr2 = runner3_params(L, par[:31])
    pv.clamp_eps, pv.clamp_lo = float(derived[0][7]), 0.0                     # F/:4458
    terms = (OFkeTerms * 2)()
    gD, gV = [], []
    for s in range(2):
        d = derived[s]
        bD, b1, b2 = int(d[4]), int(d[5]), int(d[6])
        terms[s] = OFkeTerms(L, bD, b1, b2, b1, b2, float(d[3]), float(d[7]))   # CPU path: (.first, .second) on both axes, F/:3640
        sig = float(par[15 + s])                                                  # sigma[s], F/:3769-3771
        gD.append(fke_gaussian(sig, bD)); gV.append(fke_gaussian(sig, b1))
    M = np.array([par[25], par[26]], dtype=np.float64)
    v = np.array([par[17], par[18]], dtype=np.float64)
    return pv, pc, M, v, terms, gD, gV, R, fr1, fr2


def step_fpe_3sp(pv, pc, M, v, terms, gD, gV, central, fr1, fr2, rng, f, gamma, veff, trig_libm=False, means=None):
    """One Rietkerk_FPE SPB = 3 runner step (tmo_step_fpe_3sp) in place on f [5][L*L], gamma [2][L*L] and veff [2][L*L][2]."""
    cen = np.ascontiguousarray(central, dtype=np.int32)
    m = None if means is None else np.array(means, dtype=np.float64)
    pD, pV = (dp * 2)(ptr(gD[0]), ptr(gD[1])), (dp * 2)(ptr(gV[0]), ptr(gV[1]))
    rc = oracle().tmo_step_fpe_3sp(C.byref(pv), C.byref(pc), ptr(M), ptr(v), terms, pD, pV, cen.ctypes.data_as(C.POINTER(C.c_int32)), len(cen),
                                   fr1, fr2, C.byref(rng), ptr(f), ptr(gamma), ptr(veff), 1 if trig_libm else 0, None if m is None else ptr(m))
    assert rc == 0


def fpe_runner_params(L, par, derived):
    """Oracle parameter blocks of the Rietkerk_FPE SPB = 2 runner from the vector of tests/golden/make_fpe_runner_golden.py
    ({...runner2 vector..., dtAdv1}) and the constants set_global_FKE_params derived for the grazer
    ({beta_vel, sigma_D, sigma_vD, mu_vel_prefactor, bound_D, bound_vD.first, bound_vD.second, epsilon}).
    Returns (vegetation, RK4, M1, movement terms, gD, gV, disc radius, radius fraction)."""
    pv, _, pc, R, fr = runner2_params(L, par[:23])
    pv.clamp_eps, pv.clamp_lo = float(derived[7]), 0.0                       # F/:4458
    bD, b1, b2 = int(derived[4]), int(derived[5]), int(derived[6])
    terms = OFkeTerms(L, bD, b1, b2, b1, b2, float(derived[3]), float(derived[7]))   # CPU path: (.first, .second) on both axes, F/:3640
    sigma1 = float(par[14])
    return pv, pc, float(par[19]), terms, fke_gaussian(sigma1, bD), fke_gaussian(sigma1, b1), R, fr   # stencils: F/:3769-3771


def step_fpe_2sp(pv, pc, M1, terms, gD, gV, central, fr, rng, f, gamma, veff=None, means=None):
    """One Rietkerk_FPE SPB = 2 runner step (tmo_step_fpe_2sp) in place on f [4][L*L] and gamma [L*L]."""
    cen = np.ascontiguousarray(central, dtype=np.int32)
    if veff is None:
        veff = np.zeros((pv.L * pv.L, 2))
    m = None if means is None else np.array([means[0], means[1]], dtype=np.float64)
    rc = oracle().tmo_step_fpe_2sp(C.byref(pv), C.byref(pc), M1, C.byref(terms), ptr(gD), ptr(gV), cen.ctypes.data_as(C.POINTER(C.c_int32)),
                                   len(cen), fr, C.byref(rng), ptr(f), ptr(gamma), ptr(veff), None if m is None else ptr(m))
    assert rc == 0


def dp_noise_off(L, a, b, D, dt, dx, rho, nsteps):
    """nsteps RK4 steps of percNoDiff_2D's time loop (tmo_dp_noise_off_step) on a copy of rho."""
    r = np.ascontiguousarray(rho, dtype=np.float64).copy()
    for _ in range(nsteps):
        assert oracle().tmo_dp_noise_off_step(L, a, b, D, dt, dx, ptr(r)) == 0
    return r
