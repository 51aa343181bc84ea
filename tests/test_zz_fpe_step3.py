"""The Rietkerk_FPE flavour of the vegetation + grazer + predator step (tm_cons3_set_fpe): the body of the time loop of
rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=3 (F/:4094-4462).  The oracle's tmo_step_fpe_3sp is pinned on frames of the UNMODIFIED
reference runner (tests/test_fpe_runner_oracle.py, tests/golden/runner_fpe_3sp.npz); here the product's launch chain follows it step by
step on Philox streams: CPU = the kernel SOURCE on host threads, GPU (-m gpu) = the C ABI.  Given the product's three device-side means,
vegetation, soil water, surface water, both perception fields and the advection vectors equal the oracle bit for bit; the two moved
fields agree to <= 1e-12 of their maximum (gather order vs scatter order)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
RUN3 = np.load(os.path.join(HERE, "golden", "runner_fpe_3sp.npz"))
G3 = np.load(os.path.join(HERE, "golden", "runner_3sp.npz"))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def setup(L, seed, pred_scale=1.0):
    par, derived = RUN3["mixed_par"].astype(np.float64), RUN3["mixed_derived"].astype(np.float64)
    pv, pc, M, v, terms, gD, gV, R, fr1, fr2 = O.fpe_runner3_params(L, par, derived)
    central = np.ascontiguousarray(G3["central_R%d" % R], dtype=np.int32)
    init = RUN3["mixed_init"]
    rng = np.random.default_rng(seed)
    f = np.empty((5, L * L))
    for s in range(5):
        f[s] = rng.choice(init[s], L * L)
    f[2] *= pred_scale
    return par, derived, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f


def pack(par, derived, fr1, fr2):
    (a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, D4, s0, s1, s2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2,
     dt, dx) = [float(x) for x in par[:31]]
    par3 = np.array([dt, dx, d, D0, s0, float(derived[0][7]), 0.0, D1, D2, v1, v2, m1, m2, s1, s2, D0 / (dx * dx), fr1, fr2])
    cpar = np.array([dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), aij, hij, ej, ajm, hjm, em])
    fke = np.array([[d_[4], d_[5], d_[6], d_[5], d_[6]] for d_ in derived], dtype=np.int32)
    pref = np.array([derived[0][3], derived[1][3]])
    return par3, cpar, fke, pref, np.array([m1, m2])


def emu_step(emu, L, par, derived, central, fr1, fr2, gD, gV, f, gamma, step, seed):
    par3, cpar, fke, pref, Mm = pack(par, derived, fr1, fr2)
    f, g = f.copy(), gamma.copy()
    veff, dyn, ovf = np.zeros((2, L * L, 2)), np.zeros(6), C.c_ulonglong()
    rc = emu.emu_cons_step3_fpe(L, par3.ctypes.data_as(dp), cpar.ctypes.data_as(dp), central.ctypes.data_as(C.POINTER(C.c_int)), len(central),
                                C.c_ulonglong(seed), 0, step, 1, None, None, 0, fke.ctypes.data_as(C.POINTER(C.c_int)), pref.ctypes.data_as(dp),
                                C.c_double(derived[0][7]), gD[0].ctypes.data_as(dp), gV[0].ctypes.data_as(dp), gD[1].ctypes.data_as(dp),
                                gV[1].ctypes.data_as(dp), Mm.ctypes.data_as(dp), f.ctypes.data_as(dp), g.ctypes.data_as(dp),
                                veff.ctypes.data_as(dp), dyn.ctypes.data_as(dp), C.byref(ovf))
    assert rc == 0 and ovf.value == 0
    return f, g, veff, dyn


def follow(one_step, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f, nsteps, seed):
    """one_step(f, gamma, step) -> (f', gamma', v_eff, means[6]).  The oracle restarts every step from the product's state (the movement
    step's few-ulp differences must not accumulate) with the product's three means."""
    fp, gp = f.copy(), np.zeros((2, f.shape[1]))
    for s in range(nsteps):
        fo, go, vo = fp.copy(), gp.copy(), np.zeros((2, f.shape[1], 2))
        m = [O.mean(fo[k]) for k in range(3)]
        occ = O.occupied(fo[0])
        fp, gp, vp, dyn = one_step(fp, gp, s)
        tol = max(1e-13, 8 * np.sqrt(fo[0].size) * 2.0 ** -53)
        for k in range(3):
            assert abs(dyn[k] - m[k]) <= tol * abs(m[k]) + 1e-300
        assert dyn[3] == occ
        O.step_fpe_3sp(pv, pc, M, v, terms, gD, gV, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=seed, trial=0, step=s), fo, go, vo,
                       means=(dyn[0], dyn[1], dyn[2]))
        fo = fo.reshape(fp.shape)
        for k in (0, 3, 4):
            assert np.array_equal(fp[k], fo[k]), "step %d, field %d: %d cells differ" % (s, k, np.count_nonzero(fp[k] != fo[k]))
        assert np.array_equal(gp, go), "step %d: gamma differs" % s
        assert np.array_equal(vp, vo), "step %d: advection vectors differ" % s
        for k in (1, 2):
            assert np.abs(fp[k] - fo[k]).max() <= 1e-12 * max(np.abs(fo[k]).max(), 1e-300), "step %d field %d: %.3e" % (s, k, np.abs(fp[k] - fo[k]).max())
    return fp, gp


@pytest.mark.parametrize("L,nsteps,pred_scale", [(16, 2, 1.0), (16, 2, 0.0), (150, 1, 1.0)])
def test_emulated_fpe_step3_follows_the_oracle(emu, L, nsteps, pred_scale):
    """all three species; with the predators gone (gamma_2 = 1 everywhere, F/:3430; the grazer's disc sum undiscounted); and L = 150: two
    column tiles per lattice row"""
    par, derived, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f = setup(L, 7, pred_scale)
    fp, gp = follow(lambda f_, g_, s: emu_step(emu, L, par, derived, central, fr1, fr2, gD, gV, f_, g_, s, 99),
                    pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f, nsteps, 99)
    if pred_scale == 0.0:
        assert np.all(gp[1] == 1.0) and np.all(fp[2] == 0.0)


def product_setup(L, par, derived, central, fr1, fr2, gD, gV):
    import trophic_maelstorm_b200 as tm
    (a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, D4, s0, s1, s2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2,
     dt, dx) = [float(x) for x in par[:31]]
    veg = tm.rietkerk_params(L, a=a, dt=dt)
    veg.dx, veg.rc, veg.gmax, veg.alpha, veg.rW, veg.W0, veg.d, veg.K1, veg.K2 = dx, c, gmax, alpha, rW, W0, d, K1, K2
    veg.D0, veg.D1, veg.D2, veg.sigma0, veg.clamp_eps, veg.clamp_lo = D0, D3, D4, s0, float(derived[0][7]), 0.0
    graz = tm.GrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    pred = tm.GrazParams(L, 0, dt, dx, D2, v2, m2, s2, D0 / (dx * dx))
    cons = tm.Cons3Params(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), aij, hij, ej, ajm, hjm, em)
    ops = tm.ConsumerOps3(L)
    ops.setup(veg, graz, pred, cons, central, fr1, fr2)
    terms = (tm.FkeTerms * 2)()
    for s in range(2):
        d_ = derived[s]
        terms[s] = tm.FkeTerms(int(d_[4]), int(d_[5]), int(d_[6]), int(d_[5]), int(d_[6]), len(gD[s]), len(gV[s]), 0, float(d_[3]), float(d_[7]))
    ops.set_fpe(terms, gD, gV, m1, m2)
    return ops


@pytest.mark.gpu
@pytest.mark.parametrize("L,pred_scale", [(16, 1.0), (96, 1.0), (32, 0.0), (150, 1.0)])
def test_gpu_fpe_step3_follows_the_oracle(L, pred_scale):
    par, derived, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f = setup(L, 7, pred_scale)
    ops = product_setup(L, par, derived, central, fr1, fr2, gD, gV)

    def one_step(f_, g_, s):
        ops.upload(f_, g_)
        ops.run(99, 0, s, 1)
        fn, gn, vn, means, ovf = ops.download()
        assert ovf == 0
        return fn, gn, vn, means

    follow(one_step, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f, 3, 99)
    ms, n = ops.last_ms()
    assert n == 18
    ops.close()


@pytest.mark.gpu
def test_gpu_fpe3_runner_frames_and_rows():
    """rietkerk_DorFPE_2D_MultiSp (SPB = 3 signature): the frames the runner sees at its measurement bins against the oracle stepping from
    the same Philox streams on the reference's schedule, and the 18-column rows."""
    L, r, seed = 32, 2, 321
    par, derived, pv, pc, M, v, terms, gD, gV, central, fr1, fr2, f0 = setup(L, 21)
    inits = np.stack([f0, setup(L, 22)[-1]])
    dt = pv.dt
    t_meas = np.array([0.0, 2 * dt, 4 * dt])
    h = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_fpe3_run.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double]
    h.tmh_stream_seed.restype = C.c_ulonglong
    h.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    h.tmh_set_options(0, seed, 0, None)
    key = h.tmh_stream_seed(seed, float(par[0]))
    rows = np.zeros((16, 18))
    frames = np.zeros((r, len(t_meas), 5, L * L))
    n = h.tmh_fpe3_run(L, par.ctypes.data_as(dp), r, inits.ctypes.data_as(dp), r, t_meas.ctypes.data_as(dp), len(t_meas),
                       rows.ctypes.data_as(dp), 16, frames.ctypes.data_as(dp), 0, 0, 0.0, 0.0)
    assert n == len(t_meas)
    for j in range(r):
        f, g, ve = inits[j].copy(), np.zeros((2, L * L)), np.zeros((2, L * L, 2))
        t, index, step = 0.0, 0, 0
        while t < t_meas[-1] + dt:
            if index < len(t_meas) and (t == 0 or (t >= t_meas[index] - dt / 2.0 and t < t_meas[index] + dt / 2.0)):
                got = frames[j, index]
                scale = np.abs(f).max(axis=1, keepdims=True)
                assert (np.abs(got - f) / scale).max() <= 1e-10, "replicate %d bin %d: %.3e" % (j, index, (np.abs(got - f) / scale).max())
                index += 1
            O.step_fpe_3sp(pv, pc, M, v, terms, gD, gV, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=key, trial=j, step=step), f, g, ve)
            step += 1
            t += dt
    assert np.all(rows[:n, 0] == par[0]) and np.all(rows[:n, 1] == r) and np.all(rows[:n, 2] == L) and np.allclose(rows[:n, 3], t_meas)
    for i in range(n):
        assert abs(rows[i, 4] - np.mean([O.mean(frames[j, i, 3]) for j in range(r)])) <= 1e-12 * abs(rows[i, 4])
        assert rows[i, 8] == r and rows[i, 12] == r and rows[i, 16] == r
