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
