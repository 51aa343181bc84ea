"""The SPB = 3 (vegetation + grazer + predator) time step (VERDICT r1 missing #3): vegetation Dornic -> device-side
means / cover -> the two perception fields of calc_gamma_3Sp_NonRefugia -> grazer and predator Dornic with upwind advection ->
five-field RK4 over f_2Dor_3Sp -> clamp.

The oracle's tmo_step_3sp is pinned on the UNMODIFIED reference compiled with -DSPB=3 (tests/test_consumer3_oracle.py,
tests/golden/runner_3sp.npz).  Here the product's launch chain (cons_step3_enqueue, shared by the C ABI and tests/emu) is compared with
it on per-cell streams (Philox and replay): CPU = the kernel SOURCE on host threads, GPU (-m gpu) = the C ABI.  Given the product's
three device-side means (fixed summation tree; <= max(1e-13, 8 sqrt(n) ulp) from the sequential sums) every field, both perception
fields and both sets of advection vectors equal the oracle BIT FOR BIT."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
G3 = np.load(os.path.join(HERE, "golden", "runner_3sp.npz"))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def setup(L, seed, graz_scale=1.0, pred_scale=1.0):
    par = G3["mixed_par"].astype(np.float64)
    pv, pg1, pg2, pc, R, fr1, fr2 = O.runner3_params(L, par)
    central = np.ascontiguousarray(G3["central_R%d" % R], dtype=np.int32)
    init = G3["mixed_init"]
    rng = np.random.default_rng(seed)
    f = np.empty((5, L * L))
    for s in range(5):
        f[s] = rng.choice(init[s], L * L)
    f[1] *= graz_scale
    f[2] *= pred_scale
    return par, pv, pg1, pg2, pc, central, fr1, fr2, f


def pack(par, fr1, fr2):
    (a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, D4, s0, s1, s2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2,
     dt, dx) = [float(x) for x in par]
    par3 = np.array([dt, dx, d, D0, s0, 1e-12, -1e-80, D1, D2, v1, v2, m1, m2, s1, s2, D0 / (dx * dx), fr1, fr2])
    cpar = np.array([dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), aij, hij, ej, ajm, hjm, em])
    return par3, cpar


def emu_step(emu, L, par, central, fr1, fr2, f, gamma, step, seed, u=None, stride=0):
    par3, cpar = pack(par, fr1, fr2)
    f, g = f.copy(), gamma.copy()
    veff, dyn, ovf = np.zeros((2, L * L, 2)), np.zeros(6), C.c_ulonglong()
    us = [None, None, None] if u is None else [x.ctypes.data_as(dp) for x in u]
    rc = emu.emu_cons_step3(L, par3.ctypes.data_as(dp), cpar.ctypes.data_as(dp), central.ctypes.data_as(C.POINTER(C.c_int)), len(central),
                            C.c_ulonglong(seed), 0, step, 1, us[0], us[1], us[2], stride, f.ctypes.data_as(dp), g.ctypes.data_as(dp),
                            veff.ctypes.data_as(dp), dyn.ctypes.data_as(dp), C.byref(ovf))
    assert rc == 0
    return f, g, veff, dyn, ovf.value


def follow(one_step, pv, pg1, pg2, pc, central, fr1, fr2, f, nsteps, seed):
    """one_step(f, gamma, step) -> (f', gamma', v_eff, means[6], overflow).  The oracle follows with the product's three means; the
    states must stay identical bit for bit."""
    fo, go = f.copy(), np.zeros((2, f.shape[1]))
    fp, gp = f.copy(), np.zeros((2, f.shape[1]))
    for s in range(nsteps):
        m = [O.mean(fo[k]) for k in range(3)]
        occ = O.occupied(fo[0])
        skipped = [O.mean(fo[k]) == 0.0 for k in (1, 2)]
        fp, gp, vp, dyn, ovf = one_step(fp, gp, s)
        assert ovf == 0
        tol = max(1e-13, 8 * np.sqrt(fo[0].size) * 2.0 ** -53)
        for k in range(3):
            assert abs(dyn[k] - m[k]) <= tol * abs(m[k]) + 1e-300
        assert dyn[3] == occ
        vo = O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=seed, trial=0, step=s), fo, go,
                        means=(dyn[0], dyn[1], dyn[2]))
        assert np.array_equal(fp, fo.reshape(fp.shape)), "step %d: fields differ in %d cells" % (s, np.count_nonzero(fp != fo.reshape(fp.shape)))
        assert np.array_equal(gp, go), "step %d: gamma differs" % s
        for k in (0, 1):
            if not skipped[k]:
                assert np.array_equal(vp[k], vo[k]), "step %d: advection vectors of consumer %d" % (s, k + 1)
    return fp, gp


def test_emulated_step3_follows_the_oracle(emu):
    L = 18
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 7)
    follow(lambda f_, g_, s: emu_step(emu, L, par, central, fr1, fr2, f_, g_, s, 99), pv, pg1, pg2, pc, central, fr1, fr2, f, 3, 99)


@pytest.mark.parametrize("gs,ps", [(1.0, 0.0), (0.0, 1.0), (0.0, 0.0)])
def test_emulated_step3_with_a_species_gone(emu, gs, ps):
    """no predators: the grazer's disc sum is not discounted (R/:3221) and the predator's loop is skipped (R/:3910); no grazers: both
    perception fields are zero and the grazer's loop is skipped"""
    L = 16
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 11, graz_scale=gs, pred_scale=ps)
    fp, gp = follow(lambda f_, g_, s: emu_step(emu, L, par, central, fr1, fr2, f_, g_, s, 5), pv, pg1, pg2, pc, central, fr1, fr2, f, 2, 5)
    if ps == 0.0:
        assert np.all(fp[2] == 0.0) and np.all(gp[1] == 0.0)
    if gs == 0.0:
        assert np.all(fp[1] == 0.0) and np.all(gp[0] == 0.0)


def test_emulated_step3_replay_equals_oracle_replay(emu):
    """host-supplied uniforms: cell i of species s takes its k-th uniform from u_s[i*stride + k]"""
    L, stride = 14, 48
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 13)
    rng = np.random.default_rng(2)
    u = [np.ascontiguousarray(rng.random(L * L * stride)) for _ in range(3)]
    fp, gp, vp, dyn, ovf = emu_step(emu, L, par, central, fr1, fr2, f, np.zeros((2, L * L)), 0, 0, u=u, stride=stride)
    assert ovf == 0
    fo, go = f.copy(), np.zeros((2, L * L))
    u_all = np.concatenate(u)                      # must outlive the call: the oracle's generator holds a raw pointer
    vo = O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_REPLAY, u=u_all, stride=stride), fo, go,
                    means=(dyn[0], dyn[1], dyn[2]))
    assert np.array_equal(gp, go) and np.array_equal(vp, vo)
    err = np.abs(fp - fo) / np.maximum(np.abs(fo), 1e-300)
    err[(fo == 0) & (fp == 0)] = 0
    assert err.max() <= 1e-12, (err.max(), [np.count_nonzero(err[k] > 1e-12) for k in range(5)])
    for k in (0, 1, 2):
        assert np.array_equal(fp[k] > 0, fo[k] > 0)
