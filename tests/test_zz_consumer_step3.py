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


@pytest.mark.parametrize("L,nsteps", [(16, 2), (150, 1)])   # 150: two column tiles per lattice row, the second one 22 columns wide
def test_emulated_step3_follows_the_oracle(emu, L, nsteps):
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 7)
    follow(lambda f_, g_, s: emu_step(emu, L, par, central, fr1, fr2, f_, g_, s, 99), pv, pg1, pg2, pc, central, fr1, fr2, f, nsteps, 99)


@pytest.mark.parametrize("gs,ps", [(1.0, 0.0), (0.0, 1.0)])
def test_emulated_step3_with_a_species_gone(emu, gs, ps):
    """no predators: the grazer's disc sum is not discounted (R/:3221) and the predator's loop is skipped (R/:3910); no grazers: both
    perception fields are zero and the grazer's loop is skipped"""
    L = 14
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


# ---- the C ABI (include/tm_consumer3.h) and the host runner (host/tm_host_3sp.h) on the GPU -------------------------------------------
def product_setup(L, par, central, fr1, fr2):
    import trophic_maelstorm_b200 as tm
    (a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, D4, s0, s1, s2, v1, v2, aij, hij, ej, ajm, hjm, em, m1, m2, pR1, pR2,
     dt, dx) = [float(x) for x in par]
    veg = tm.rietkerk_params(L, a=a, dt=dt)
    veg.dx, veg.rc, veg.gmax, veg.alpha, veg.rW, veg.W0, veg.d, veg.K1, veg.K2 = dx, c, gmax, alpha, rW, W0, d, K1, K2
    veg.D0, veg.D1, veg.D2, veg.sigma0, veg.clamp_eps, veg.clamp_lo = D0, D3, D4, s0, 1e-12, -1e-80
    graz = tm.GrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    pred = tm.GrazParams(L, 0, dt, dx, D2, v2, m2, s2, D0 / (dx * dx))
    cons = tm.Cons3Params(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), aij, hij, ej, ajm, hjm, em)
    ops = tm.ConsumerOps3(L)
    ops.setup(veg, graz, pred, cons, central, fr1, fr2)
    return ops


@pytest.mark.gpu
@pytest.mark.parametrize("L,gs,ps", [(18, 1.0, 1.0), (96, 1.0, 1.0), (32, 1.0, 0.0), (32, 0.0, 1.0), (150, 1.0, 1.0)])
def test_gpu_step3_follows_the_oracle(L, gs, ps):
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 7, graz_scale=gs, pred_scale=ps)
    ops = product_setup(L, par, central, fr1, fr2)

    def one_step(f_, g_, s):
        ops.upload(f_, g_)
        ops.run(99, 0, s, 1)
        return ops.download()

    follow(one_step, pv, pg1, pg2, pc, central, fr1, fr2, f, 3, 99)
    ms, n = ops.last_ms()
    assert n == 10
    ops.close()


@pytest.mark.gpu
def test_gpu_step3_replay_and_resident_steps():
    L, stride = 24, 48
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = setup(L, 13)
    ops = product_setup(L, par, central, fr1, fr2)
    rng = np.random.default_rng(2)
    u = [np.ascontiguousarray(rng.random(L * L * stride)) for _ in range(3)]
    ops.upload(f, None)
    ops.replay(u[0], u[1], u[2], stride)
    fp, gp, vp, dyn, ovf = ops.download()
    assert ovf == 0
    fo, go = f.copy(), np.zeros((2, L * L))
    u_all = np.concatenate(u)
    vo = O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_REPLAY, u=u_all, stride=stride), fo, go, means=(dyn[0], dyn[1], dyn[2]))
    assert np.array_equal(gp, go) and np.array_equal(vp, vo)
    err = np.abs(fp - fo) / np.maximum(np.abs(fo), 1e-300)
    err[(fo == 0) & (fp == 0)] = 0
    assert err.max() < 1e-6 and np.count_nonzero(err[:3] > 1e-12) <= 3      # at most a Gaussian-truncation flip or two
    # several steps enqueued at once = the same steps one at a time
    ops.upload(f, None); ops.run(17, 0, 0, 4); a = ops.download()
    ops.upload(f, None)
    for s in range(4):
        ops.run(17, 0, s, 1)
    b = ops.download()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.isfinite(a[0]).all()
    ops.close()


@pytest.mark.gpu
def test_gpu_spb3_runner_frames_and_rows():
    """rietkerk_Dornic_2D_MultiSp (SPB = 3 signature) over the resident stepper: the frames it measures against the oracle stepped from
    the same Philox streams (seed = stream_seed(options().seed, a), trial = replicate), and the 18-column rows."""
    L, r, seed = 24, 2, 4242
    par, pv, pg1, pg2, pc, central, fr1, fr2, _ = setup(L, 1)
    inits = np.stack([setup(L, 50 + j)[8] for j in range(r)])
    t_meas = np.array([0.0, 0.2, 0.5])
    h = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_riet3_run.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double]
    h.tmh_stream_seed.restype = C.c_ulonglong
    h.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    h.tmh_set_options(0, seed, 0, None)
    key = h.tmh_stream_seed(seed, float(par[0]))
    rows = np.zeros((16, 18))
    frames = np.zeros((r, len(t_meas), 5, L * L))
    n = h.tmh_riet3_run(L, par.ctypes.data_as(dp), r, inits.ctypes.data_as(dp), r, t_meas.ctypes.data_as(dp), len(t_meas),
                        rows.ctypes.data_as(dp), 16, frames.ctypes.data_as(dp), 0, 0, 0.0, 0.0)
    assert n == len(t_meas)
    dt = pv.dt
    for j in range(r):
        f, g = inits[j].copy(), np.zeros((2, L * L))
        t, index, step = 0.0, 0, 0
        while t < t_meas[-1] + dt:
            if index < len(t_meas) and (t == 0 or (t >= t_meas[index] - dt / 2.0 and t < t_meas[index] + dt / 2.0)):
                got = frames[j, index]
                scale = np.abs(f).max(axis=1, keepdims=True)
                assert (np.abs(got - f) / scale).max() <= 1e-9, "replicate %d bin %d: %.3e" % (j, index, (np.abs(got - f) / scale).max())
                index += 1
            O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=key, trial=j, step=step), f, g)
            step += 1
            t += dt
    assert np.all(rows[:n, 0] == par[0]) and np.all(rows[:n, 1] == r) and np.all(rows[:n, 2] == L) and np.allclose(rows[:n, 3], t_meas)
    for i in range(n):
        assert abs(rows[i, 4] - np.mean([O.mean(frames[j, i, 3]) for j in range(r)])) <= 1e-12 * abs(rows[i, 4])
        assert abs(rows[i, 5] - np.mean([O.mean(frames[j, i, 4]) for j in range(r)])) <= 1e-12 * abs(rows[i, 5])
        assert rows[i, 8] == r and rows[i, 12] == r and rows[i, 16] == r
