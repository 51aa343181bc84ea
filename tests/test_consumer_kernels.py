"""Consumer-side kernels (csrc/tm_consumer.cuh, SURVEY 8 f3 first part): perception field and four-field RK4.

CPU: the product's kernel SOURCE on host threads (tests/emu) against the oracle (bit for bit) and the reference's goldens.
GPU (-m gpu): the same through the C ABI (include/tm_consumer.h), plus larger lattices."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "consumer_ops.npz"))
dp = C.POINTER(C.c_double)
GAMMA_CASES = ["gamma_full", "gamma_sparse", "gamma_small", "gamma_dead"]


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def gamma_case(name):
    L, R, fr, nveg = GOLD[name + "_cfg"]
    central = GOLD["central_R%d" % int(R)]
    new_length, eff = O.gamma_window(len(central), float(fr), float(nveg))
    avg = GOLD[name + "_avg"]
    return int(L), np.ascontiguousarray(central[:new_length][:eff], dtype=np.int32), GOLD[name + "_veg"], float(avg[0]), float(avg[1]), \
        GOLD[name + "_g0"], GOLD[name + "_out"]


def tm_params(L, par):
    o = O.cons_params(L, par)
    return tm.ConsParams(*[getattr(o, n) for n, _ in tm.ConsParams._fields_])


def random_fields(L, seed):
    rng = np.random.default_rng(seed)
    P = rng.gamma(2.0, 3000.0, L * L); P[rng.random(L * L) < 0.2] = 0.0
    G = rng.gamma(2.0, 2.0, L * L); G[rng.random(L * L) < 0.3] = 0.0
    W = 0.25 * (1 + 0.2 * rng.standard_normal(L * L)); O_ = 6.6 * (1 + 0.2 * rng.standard_normal(L * L))
    return np.ascontiguousarray(np.stack([P, G, W, O_]))


@pytest.mark.parametrize("name", GAMMA_CASES)
def test_emulated_gamma_matches_reference(emu, name):
    L, off, veg, a0, a1, g0, want = gamma_case(name)
    g = g0.copy()
    emu.emu_cons_gamma(L, veg.ctypes.data_as(dp), off.ctypes.data_as(C.POINTER(C.c_int)), len(off), C.c_double(a0), C.c_double(a1), g.ctypes.data_as(dp))
    assert np.array_equal(g, want)


@pytest.mark.parametrize("name", ["rk4_a", "rk4_b"])
def test_emulated_rk4_matches_oracle_and_reference(emu, name):
    L, steps = [int(x) for x in GOLD[name + "_cfg"]]
    par = GOLD["rk4_par"]
    o = O.cons_params(L, par)
    flat = np.array([getattr(o, n) for n in "dt a c gmax alpha rW W0 K1 K2 Dw Do A01 H01 E1".split()])
    f = GOLD[name + "_in"].copy()
    emu.emu_cons_rk4(L, flat.ctypes.data_as(dp), f.ctypes.data_as(dp), steps)
    assert np.array_equal(f, O.rk4_2sp(o, GOLD[name + "_in"], steps))          # same operations in the same order
    want = GOLD[name + "_out"]
    assert (np.abs(f - want) / np.abs(want).max(axis=1, keepdims=True)).max() <= 1e-12


def test_abi_exports_consumer_symbols():
    lib = tm.lib()
    for name in tm.CONS_ABI_SYMBOLS:
        getattr(lib, name)
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "tm_consumer.h")).read()
    declared = set(re.findall(r"^(?:int|void \*|const char \*)\s*(tm_cons_[a-z0-9_]+)\(", hdr, re.M))
    assert declared == set(tm.CONS_ABI_SYMBOLS)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", GAMMA_CASES)
def test_gpu_gamma_matches_reference(name):
    L, off, veg, a0, a1, g0, want = gamma_case(name)
    got = tm.ConsumerOps(L).gamma(veg, off, a0, a1, g0)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rk4_a", "rk4_b"])
def test_gpu_rk4_matches_oracle_and_reference(name):
    L, steps = [int(x) for x in GOLD[name + "_cfg"]]
    par = GOLD["rk4_par"]
    got = tm.ConsumerOps(L).rk4(tm_params(L, par), GOLD[name + "_in"], steps)
    assert np.array_equal(got, O.rk4_2sp(O.cons_params(L, par), GOLD[name + "_in"], steps))
    want = GOLD[name + "_out"]
    assert (np.abs(got - want) / np.abs(want).max(axis=1, keepdims=True)).max() <= 1e-12


@pytest.mark.gpu
def test_gpu_consumer_ops_larger_lattices():
    """L = 300 against the oracle (bit for bit); L = 2048: gamma in [0, 1] and = 1/eff-normalised disc mean for a flat
    field; the grazer-free RK4 leaves G = 0 and equals the SPB = 1 RK4 of the lattice oracle (f_2Dor, R/:2389)."""
    par = GOLD["rk4_par"]
    L = 300
    f0 = random_fields(L, 5)
    ops = tm.ConsumerOps(L)
    assert np.array_equal(ops.rk4(tm_params(L, par), f0, 3), O.rk4_2sp(O.cons_params(L, par), f0, 3))
    central = GOLD["central_R5"]
    g0 = np.random.default_rng(1).random(L * L)
    want = O.gamma_perception(L, f0[0], central, len(central), f0[0].mean(), f0[1].mean(), g0)
    assert np.array_equal(ops.gamma(f0[0], central, f0[0].mean(), f0[1].mean(), g0), want)
    # grazer-free: the (P, W, O) part must equal the three-field RK4 of the SPB = 1 oracle bit for bit
    f1 = f0.copy(); f1[1] = 0.0
    got = ops.rk4(tm_params(L, par), f1, 2)
    p1 = tm.rietkerk_params(L)
    three = np.ascontiguousarray(f1[[0, 2, 3]])
    for _ in range(2):
        assert O.oracle().tmo_rk4(C.byref(O.to_oparams(p1)), O.ptr(three), 0) == 0
    assert np.all(got[1] == 0.0) and np.array_equal(got[[0, 2, 3]], three)
    L = 2048
    ops = tm.ConsumerOps(L)
    flat = np.full(L * L, 123.0)
    g = ops.gamma(flat, central, 123.0, 1.0, np.zeros(L * L))
    assert np.allclose(g, 1.0, rtol=0, atol=1e-13)
    ms, n = ops.last_ms()
    assert n == 1 and ms > 0
    rnd = np.random.default_rng(2).gamma(2.0, 50.0, L * L)
    g = ops.gamma(rnd, central, rnd.mean(), 1.0, np.zeros(L * L))
    assert g.min() >= 0.0 and g.max() <= 1.0 and 0.9 < g.mean() <= 1.0     # disc means scatter around 1; the upper half is clamped


# ------------------------------------------------------------------------------------------------ grazer Dornic (R/:3908-4114)
def grazer_setup(L, seed, sigma=1.0):
    """Constants of the SPB = 2 driver (order_2stoc_test_rietkerk.cpp: D[1] = 0.01028, v[1] = 0.45427, 20 kg grazer) and a field
    that exercises every sampler regime: empty regions (skip rule), a thin fringe (exact Poisson/Gamma), dense bulk (Gaussian)."""
    rng = np.random.default_rng(seed)
    dx, dt = 0.1, 0.1
    M = 0.061609 * 20.0 ** -0.25 / 8760.0
    par = dict(dt=dt, dx=dx, D=0.01028, v=0.45427, M=M, sigma=sigma, D0floor=(0.00025 / 24.0) / (dx * dx))
    prev = rng.gamma(2.0, 400.0, (L, L))
    prev[:, : L // 4] = 0.0
    prev[:, L // 4: L // 4 + 3] = 10 ** rng.uniform(-4, 0, (L, 3))
    prev[rng.random((L, L)) < 0.05] = 0.0
    gamma = rng.random(L * L) * 1.1                      # some above 1: clamped (R/:3915)
    gamma[rng.random(L * L) < 0.1] = 0.0
    return par, prev.ravel().copy(), gamma


def oparams(L, par):
    return O.OGrazParams(L, 0, *[par[k] for k in "dt dx D v M sigma D0floor".split()])


def gparams(L, par):
    return tm.GrazParams(L, 0, *[par[k] for k in "dt dx D v M sigma D0floor".split()])


def emu_grazer(emu, L, par, prev, gamma, seed=0, trial=0, step=0, replay=None, stride=0):
    flat = np.array([par[k] for k in "dt dx D v M sigma D0floor".split()])
    g = gamma.copy(); out = np.empty(L * L); veff = np.empty((L * L, 2)); ovf = C.c_ulonglong()
    emu.emu_grazer_dornic(L, flat.ctypes.data_as(dp), C.c_ulonglong(seed), trial, step,
                          replay.ctypes.data_as(dp) if replay is not None else None, stride,
                          prev.ctypes.data_as(dp), g.ctypes.data_as(dp), out.ctypes.data_as(dp), veff.ctypes.data_as(dp), C.byref(ovf))
    return out, g, veff, ovf.value


@pytest.mark.parametrize("L,sigma", [(40, 1.0), (150, 0.3)])
def test_emulated_grazer_dornic_matches_oracle(emu, L, sigma):
    par, prev, gamma = grazer_setup(L, 11 + L, sigma)
    want = O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_PHILOX, seed=99, trial=2, step=7), prev, gamma)
    got = emu_grazer(emu, L, par, prev, gamma, seed=99, trial=2, step=7)
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)
    assert got[3] == 0
    stride = 48
    u = np.random.default_rng(5).random(L * L * stride)
    want = O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_REPLAY, u=u, stride=stride), prev, gamma)
    got = emu_grazer(emu, L, par, prev, gamma, replay=u, stride=stride)
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)


def test_grazer_dornic_oracle_laws():
    """Structure of the restatement: empty neighbourhoods are skipped, |v_eff| = v, and with advection switched off
    (gamma = 1) an ensemble mean follows the linear part exp(beta dt) rho + diffusion of the Dornic scheme."""
    L = 64
    par, prev, gamma = grazer_setup(L, 3)
    out, g, veff = O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_PHILOX, seed=1), prev, gamma)
    assert g.max() <= 1.0 and np.allclose(np.hypot(veff[:, 0], veff[:, 1]), par["v"], rtol=1e-14)
    P = prev.reshape(L, L)
    nb = np.roll(P, 1, 0) + np.roll(P, -1, 0) + np.roll(P, 1, 1) + np.roll(P, -1, 1)
    dead = (P == 0) & (nb == 0)
    assert dead.sum() > 100 and np.all(out.reshape(L, L)[dead] == 0.0)
    assert np.all(out >= 0) and np.isfinite(out).all()
    flat = np.full(L * L, 500.0)
    outs = [O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_PHILOX, seed=s), flat, np.ones(L * L))[0].mean() for s in range(8)]
    # homogeneous field, gamma = 1: E[rho*] = rho e^{-M dt} (diffusion in = diffusion out); sampling noise ~ sigma sqrt(rho dt)/L
    assert abs(np.mean(outs) - 500.0 * np.exp(-par["M"] * par["dt"])) < 5 * np.sqrt(500.0 * par["dt"]) / L / np.sqrt(8) + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("L,sigma", [(40, 1.0), (150, 0.3), (512, 1.0)])
def test_gpu_grazer_dornic_matches_oracle(L, sigma):
    par, prev, gamma = grazer_setup(L, 11 + L, sigma)
    ops = tm.ConsumerOps(L)
    want = O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_PHILOX, seed=99, trial=2, step=7), prev, gamma)
    got = ops.grazer_dornic(gparams(L, par), prev, gamma, seed=99, trial=2, step=7)
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)
    stride = 48
    u = np.random.default_rng(5).random(L * L * stride)
    want = O.grazer_dornic(oparams(L, par), O.make_rng(O.SRC_REPLAY, u=u, stride=stride), prev, gamma)
    got = ops.grazer_dornic(gparams(L, par), prev, gamma, replay=u, stride=stride)
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)
    assert got[3] == 0
    short = ops.grazer_dornic(gparams(L, par), prev, gamma, replay=u[: L * L * 2].copy(), stride=2)
    assert short[3] > 0                               # a starved replay slice is counted, not hung
