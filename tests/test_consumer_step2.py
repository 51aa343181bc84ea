"""The SPB = 2 (vegetation + grazer) time step resident on the device (tm_cons_step2_*): vegetation Dornic ->
device-side means/cover -> perception field -> grazer Dornic with upwind advection -> four-field RK4 -> clamp.

The oracle's tmo_step_2sp is pinned on the reference's rietkerk_Dornic_2D_MultiSp compiled with -DSPB=2
(tests/test_consumer_oracle.py, tests/golden/runner_2sp.npz).  Here the product's launch chain is compared with it on
per-cell streams (Philox and replay): CPU = the kernel SOURCE on host threads (tests/emu), GPU (-m gpu) = the C ABI.
Parity statement, in two parts: (1) the two spatial means <P>, <G> the step derives on the device come from a fixed
summation tree instead of the reference's left-to-right sum and agree with it to <= max(1e-13, 8 sqrt(n) ulp) -- the
sequential sum's own rounding grows like sqrt(n); (2) GIVEN those two scalars
(tmo_step_2sp_means) every field, gamma and the advection vectors equal the oracle BIT FOR BIT, step after step.  The
free-running comparison (oracle with its own sums) is kept at small sizes (<= 1e-12): the grazer's advection-corrected
diffusion D - dx/2 (1-gamma)|v|(...) (R/:4008) cancels in some cells and amplifies the few-ulp difference of gamma."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "consumer_ops.npz"))
RUN2 = np.load(os.path.join(HERE, "golden", "runner_2sp.npz"))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def setup(L, seed, graz_scale=1.0):
    """Runner constants of the SPB = 2 golden (the driver's), an L x L frame in the golden's regime."""
    par = RUN2["mixed_par"].astype(np.float64)
    pv, pg, pc, R, fr = O.runner2_params(L, par)
    central = np.ascontiguousarray(GOLD["central_R%d" % R], dtype=np.int32)
    init = RUN2["mixed_init"]
    rng = np.random.default_rng(seed)
    f = np.empty((4, L * L))
    for s in range(4):
        f[s] = rng.choice(init[s], L * L)
    f[1] *= graz_scale
    return par, pv, pg, pc, central, fr, f


def product_params(L, par):
    a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, s0, s1, v1, aij, hij, ej, m1, pR1, dt, dx = [float(x) for x in par]
    veg = tm.rietkerk_params(L, a=a, dt=dt)
    veg.dx, veg.rc, veg.gmax, veg.alpha, veg.rW, veg.W0, veg.d, veg.K1, veg.K2 = dx, c, gmax, alpha, rW, W0, d, K1, K2
    veg.D0, veg.D1, veg.D2, veg.sigma0, veg.clamp_eps, veg.clamp_lo = D0, D2, D3, s0, 1e-12, -1e-80
    graz = tm.GrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    cons = tm.ConsParams(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D2 / (dx * dx), D3 / (dx * dx), aij, hij, ej)
    return veg, graz, cons


def emu_step2(emu, L, par, central, fr, f, gamma, nsteps=1, seed=0, trial=0, step0=0, u_veg=None, u_graz=None, stride=0, rk4_stages=0):
    a, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, s0, s1, v1, aij, hij, ej, m1, pR1, dt, dx = [float(x) for x in par]
    par2 = np.array([dt, dx, d, D0, s0, 1e-12, -1e-80, D1, v1, m1, s1, D0 / (dx * dx), fr])
    cpar = np.array([dt, a, c, gmax, alpha, rW, W0, K1, K2, D2 / (dx * dx), D3 / (dx * dx), aij, hij, ej])
    f, g = f.copy(), gamma.copy()
    veff, dyn, ovf = np.zeros((L * L, 2)), np.zeros(4), C.c_ulonglong()
    emu.emu_cons_step2(L, par2.ctypes.data_as(dp), cpar.ctypes.data_as(dp), central.ctypes.data_as(C.POINTER(C.c_int)), len(central),
                       C.c_ulonglong(seed), trial, step0, nsteps, None if u_veg is None else u_veg.ctypes.data_as(dp),
                       None if u_graz is None else u_graz.ctypes.data_as(dp), stride, f.ctypes.data_as(dp), g.ctypes.data_as(dp),
                       veff.ctypes.data_as(dp), dyn.ctypes.data_as(dp), C.byref(ovf), rk4_stages)
    return f, g, veff, dyn, ovf.value


def oracle_steps(pv, pg, pc, central, fr, f, gamma, nsteps, seed=0, trial=0, step0=0, u=None, stride=0):
    f, g = f.copy(), gamma.copy()
    veff = None
    for s in range(nsteps):
        rng = O.make_rng(O.SRC_PHILOX, seed=seed, trial=trial, step=step0 + s) if u is None else O.make_rng(O.SRC_REPLAY, u=u, stride=stride)
        veff = O.step_2sp(pv, pg, pc, central, fr, rng, f, g)
        assert rng.overflow == 0
    return f, g, veff


def stepwise_bit_parity(one_step, pv, pg, pc, central, fr, f, gamma, nsteps, seed=0, trial=0, step0=0):
    """one_step(f, gamma, step) -> (f', gamma', v_eff, means[4]) runs ONE product step.  The oracle follows with the product's
    two means; states must stay identical bit for bit."""
    fo, go = f.copy(), gamma.copy()
    fp, gp = f.copy(), gamma.copy()
    for s in range(nsteps):
        mP, mG = O.mean(fo[0]), O.mean(fo[1])
        occ = O.occupied(fo[0])
        fp, gp, vp, dyn = one_step(fp, gp, step0 + s)
        # a left-to-right sum of n terms carries ~sqrt(n) ulp of its own rounding; the tree sum is the more accurate of the two
        tol = max(1e-13, 8 * np.sqrt(fo[0].size) * 2.0 ** -53)
        assert abs(dyn[0] - mP) <= tol * abs(mP) and abs(dyn[1] - mG) <= tol * abs(mG) and dyn[2] == occ
        vo = O.step_2sp(pv, pg, pc, central, fr, O.make_rng(O.SRC_PHILOX, seed=seed, trial=trial, step=step0 + s), fo, go, means=(dyn[0], dyn[1]))
        assert np.array_equal(fp, fo.reshape(fp.shape)), "step %d: fields differ in %d cells" % (s, np.count_nonzero(fp != fo.reshape(fp.shape)))
        assert np.array_equal(gp, go) and np.array_equal(vp, vo)


def check(got, want, what, flips=0):
    """flips: cells allowed beyond 1e-12.  The tree-summed <P> is a few ulp from the sequential sum, gamma inherits that, and
    where the grazer's Gaussian Poisson shortcut truncates `long(norm*sqrt(z)+z)` (R/:3776 form) right at an integer, one
    ulp flips the count: a 1/z ~ 1e-10 relative jump in that cell (the same effect as FAST mode's)."""
    gf, gg, gv = got
    wf, wg, wv = want
    for s in (0, 1):
        assert np.array_equal(gf[s] > 0, wf[s] > 0), "%s: occupancy of field %d" % (what, s)
    scale = np.maximum(np.abs(wf), 1e-300)
    err = np.abs(gf - wf) / scale
    err[(wf == 0) & (gf == 0)] = 0
    assert np.count_nonzero(err > 1e-12) <= flips and err.max() < 1e-7, "%s: fields %.3e, %d cells beyond 1e-12" % (what, err.max(), np.count_nonzero(err > 1e-12))
    assert np.abs(gg - wg).max() <= 1e-13, "%s: gamma" % what
    assert np.array_equal(gv, wv), "%s: advection vectors" % what


@pytest.mark.parametrize("L,nsteps", [(20, 3), (48, 2)])
def test_emulated_step2_philox_matches_oracle(emu, L, nsteps):
    par, pv, pg, pc, central, fr, f = setup(L, 5 + L)
    gamma = np.zeros(L * L)
    want = oracle_steps(pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)
    got = emu_step2(emu, L, par, central, fr, f, gamma, nsteps=nsteps, seed=77, trial=3, step0=10)
    check(got[:3], want, "philox L=%d" % L)
    assert got[4] == 0

    def one(fp, gp, step):
        r = emu_step2(emu, L, par, central, fr, fp, gp, nsteps=1, seed=77, trial=3, step0=step)
        return r[0], r[1], r[2], r[3]
    if L <= 20:
        stepwise_bit_parity(one, pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)


def test_emulated_step2_replay_matches_oracle(emu):
    L, stride = 24, 64
    par, pv, pg, pc, central, fr, f = setup(L, 9)
    gamma = np.random.default_rng(2).random(L * L) * 0.5
    u = np.random.default_rng(4).random(2 * L * L * stride)
    want = oracle_steps(pv, pg, pc, central, fr, f, gamma, 1, u=u, stride=stride)
    got = emu_step2(emu, L, par, central, fr, f, gamma, u_veg=u[: L * L * stride], u_graz=u[L * L * stride:], stride=stride)
    check(got[:3], want, "replay")
    assert got[4] == 0


def test_emulated_step2_without_grazers_skips_their_loop(emu):
    """<G> == 0: the grazer's loop is skipped (R/:3910) -- G stays zero, nothing is drawn for it, gamma is zeroed."""
    L = 20
    par, pv, pg, pc, central, fr, f = setup(L, 1, graz_scale=0.0)
    gamma = np.full(L * L, 0.25)
    want = oracle_steps(pv, pg, pc, central, fr, f, gamma, 2, seed=5)
    got = emu_step2(emu, L, par, central, fr, f, gamma, nsteps=2, seed=5)
    check((got[0], got[1], want[2]), want, "no grazers")
    assert not got[0][1].any() and not got[1].any() and not got[2].any()


def test_step2_device_scalars_follow_the_reference_rules(emu):
    """The device-side scalars: <P>, <G> to a few ulp of the sequential sums, cover count exact, disc size by R/:2539, 2569."""
    L = 32
    par, pv, pg, pc, central, fr, f = setup(L, 3)
    got = emu_step2(emu, L, par, central, fr, f, np.zeros(L * L), nsteps=1, seed=1)
    mP, mG, occ, eff = got[3]
    assert abs(mP - O.mean(f[0])) <= 1e-13 * abs(mP) and abs(mG - O.mean(f[1])) <= 1e-13 * abs(mG)
    assert occ == O.occupied(f[0])
    nfac = max(1 - occ / (L * L), 0.35)
    assert eff == O.gamma_window(len(central), fr, nfac)[1]


@pytest.mark.gpu
@pytest.mark.parametrize("L,nsteps", [(20, 3), (150, 3), (512, 2)])
def test_gpu_step2_philox_matches_oracle(L, nsteps):
    par, pv, pg, pc, central, fr, f = setup(L, 5 + L)
    gamma = np.zeros(L * L)
    want = oracle_steps(pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)
    ops = tm.ConsumerOps(L)
    ops.step2_setup(*product_params(L, par), central, fr)
    ops.step2_upload(f, gamma)
    ops.step2_run(77, trial=3, step0=10, nsteps=nsteps)
    gf, gg, gv, means, ovf = ops.step2_download()
    if L < 256:
        check((gf, gg, gv), want, "gpu philox L=%d" % L)
    else:   # free-running at size: cells where R/:4008 cancels amplify the few-ulp difference of the means
        err = np.abs(gf - want[0]) / np.maximum(np.abs(want[0]), 1e-300)
        err[(want[0] == 0) & (gf == 0)] = 0
        assert np.mean(err <= 1e-12) > 0.999 and err.max() < 1e-8
    assert ovf == 0
    ms, launches = ops.last_ms()
    assert launches == 6 * nsteps and ms > 0

    def one(fp, gp, step):
        ops.step2_upload(fp, gp)
        ops.step2_run(77, trial=3, step0=step, nsteps=1)
        r = ops.step2_download()
        return r[0], r[1], r[2], r[3]
    stepwise_bit_parity(one, pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)
    ops.close()


@pytest.mark.gpu
def test_gpu_step2_replay_and_empty_grazers():
    L, stride = 64, 64
    par, pv, pg, pc, central, fr, f = setup(L, 9)
    gamma = np.random.default_rng(2).random(L * L) * 0.5
    u = np.random.default_rng(4).random(2 * L * L * stride)
    want = oracle_steps(pv, pg, pc, central, fr, f, gamma, 1, u=u, stride=stride)
    ops = tm.ConsumerOps(L)
    ops.step2_setup(*product_params(L, par), central, fr)
    ops.step2_upload(f, gamma)
    ops.step2_replay(u[: L * L * stride], u[L * L * stride:], stride)
    gf, gg, gv, means, ovf = ops.step2_download()
    check((gf, gg, gv), want, "gpu replay")
    assert ovf == 0
    f0 = f.copy(); f0[1] = 0.0
    want = oracle_steps(pv, pg, pc, central, fr, f0, gamma, 2, seed=5)
    ops.step2_upload(f0, gamma)
    ops.step2_run(5, nsteps=2)
    gf, gg, gv, means, ovf = ops.step2_download()
    check((gf, gg, want[2]), want, "gpu no grazers")
    assert not gf[1].any() and not gg.any() and means[1] == 0.0
    ops.close()


@pytest.mark.gpu
def test_gpu_step2_matches_emulation_and_is_reproducible():
    """Two runs of the resident stepper give identical bits (fixed summation tree, counter-based streams)."""
    L = 256
    par, pv, pg, pc, central, fr, f = setup(L, 12)
    ops = tm.ConsumerOps(L)
    ops.step2_setup(*product_params(L, par), central, fr)
    outs = []
    for _ in range(2):
        ops.step2_upload(f, None)
        ops.step2_run(123, nsteps=5)
        outs.append(ops.step2_download())
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.isfinite(outs[0][0]).all() and (outs[0][0] >= 0).all()
    ops.close()


# ---- the reference-facing SPB = 2 runner (host/tm_host_2sp.h) ---------------------------------------------------------
def host_lib():
    h = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_riet2_run.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp]
    h.tmh_central_sites.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    return h


def test_host_disc_list_equals_the_reference():
    """computeNeighboringSitesCentral of the host layer (std::sort with the reference's comparator) gives the reference's
    own list, ties included (golden made by the compiled reference)."""
    h = host_lib()
    n_checked = 0
    for key in GOLD.files:
        if not key.startswith("central_R"):
            continue
        R = int(key[len("central_R"):])
        want = GOLD[key]
        buf = np.zeros((len(want) + 8, 2), dtype=np.int32)
        n = h.tmh_central_sites(R, buf.ctypes.data_as(C.POINTER(C.c_int)), len(buf))
        assert n == len(want) and np.array_equal(buf[:n], want)
        n_checked += 1
    assert n_checked >= 1


def running_table(rho_m, t_meas):
    """var_mean_incremental_surv_runs R/:1293-1343 + the first-replicate rule R/:4294-4316 in plain Python (int 1/(r-1))."""
    tot = len(t_meas)
    T = np.zeros((tot, 11))
    T[:, 0] = t_meas
    for j, X in enumerate(rho_m):
        if j == 0:
            for i in range(tot):
                for s in range(2):
                    T[i, 4 * s + 3], T[i, 4 * s + 4], T[i, 4 * s + 6] = X[i, 2 * s + 1], 0.0, X[i, 2 * s]
                    T[i, 4 * s + 5] = 1 if X[i, 2 * s] > 0 else 0
                T[i, 1], T[i, 2] = X[i, 5], X[i, 7]
            continue
        for s in range(2):
            for i in range(tot):
                r = int(T[i, 4 * s + 5] + 1)
                if X[i, 2 * s] == 0.0:
                    break
                prev = T[i, 4 * s + 3]
                if T[i, 4 * s + 5] == 0.0:
                    T[i, 4 * s + 3], T[i, 4 * s + 4], T[i, 4 * s + 6] = X[i, 2 * s + 1], 0.0, X[i, 2 * s]
                else:
                    T[i, 4 * s + 3] = (X[i, 2 * s + 1] + (r - 1) * T[i, 4 * s + 3]) / r
                    T[i, 4 * s + 6] = (X[i, 2 * s] + (r - 1) * T[i, 4 * s + 6]) / r
                    T[i, 4 * s + 4] = (T[i, 4 * s + 3] - prev) ** 2 + (1 // (r - 1)) * ((r - 2) * T[i, 4 * s + 4] + (X[i, 2 * s + 1] - T[i, 4 * s + 3]) ** 2)
                T[i, 4 * s + 5] = r
        for i in range(tot):
            T[i, 1] = (X[i, 5] + j * T[i, 1]) / (j + 1)
            T[i, 2] = (X[i, 7] + j * T[i, 2]) / (j + 1)
    return T


@pytest.mark.gpu
def test_gpu_spb2_runner_rows_and_frames():
    """rietkerk_Dornic_2D_MultiSp (SPB = 2 signature) over the resident stepper: the frames it measures equal the oracle's
    stepped from the same Philox streams (seed = stream_seed(options().seed, a), trial = replicate, step = steps since t = 0), and its
    output rows are the reference's bookkeeping of those frames (left-to-right means, running mean/variance over
    surviving replicates with the integer 1/(r-1), row layout R/:4361-4372)."""
    L, r, seed = 24, 3, 4242
    par, pv, pg, pc, central, fr, _ = setup(L, 1)
    inits = np.stack([setup(L, 50 + j)[6] for j in range(r)])
    t_meas = np.array([0.0, 0.2, 0.5, 0.9, 1.0])
    h = host_lib()
    h.tmh_set_options(0, seed, 0, None)
    h.tmh_stream_seed.restype = C.c_ulonglong
    h.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    key = h.tmh_stream_seed(seed, float(par[0]))       # the runner keys its streams by (TM_SEED, a): tm_host.h stream_seed
    rows = np.zeros((16, 14))
    frames = np.zeros((r, len(t_meas), 4, L * L))
    n = h.tmh_riet2_run(L, par.ctypes.data_as(dp), r, inits.ctypes.data_as(dp), r, t_meas.ctypes.data_as(dp), len(t_meas),
                        rows.ctypes.data_as(dp), 16, frames.ctypes.data_as(dp))
    assert n == len(t_meas)
    # frames: the oracle from the same streams, measured on the reference's schedule (accumulated t += dt, R/:3648-3659)
    dt = pv.dt
    rho_m = []
    for j in range(r):
        f, g = inits[j].copy(), np.zeros(L * L)
        t, index, step = 0.0, 0, 0
        X = np.zeros((len(t_meas), 8))
        while t < t_meas[-1] + dt:
            if index < len(t_meas) and (t == 0 or (t >= t_meas[index] - dt / 2.0 and t < t_meas[index] + dt / 2.0)):
                got = frames[j, index]
                for s_ in (0, 1):
                    assert np.array_equal(got[s_] > 0, f[s_] > 0)
                err = np.abs(got - f) / np.maximum(np.abs(f), 1e-300)
                err[(f == 0) & (got == 0)] = 0
                assert err.max() <= 1e-12, "replicate %d bin %d: %.3e" % (j, index, err.max())
                for s_ in range(4):
                    X[index, 2 * s_], X[index, 2 * s_ + 1] = O.occupied(got[s_]), O.mean(got[s_])
                index += 1
            O.step_2sp(pv, pg, pc, central, fr, O.make_rng(O.SRC_PHILOX, seed=key, trial=j, step=step), f, g)
            step += 1
            t += dt
        rho_m.append(X)
    T = running_table(rho_m, t_meas)
    want = np.zeros((len(t_meas), 14))
    want[:, 0], want[:, 1], want[:, 2] = par[0], r, L
    want[:, 3:6] = T[:, 0:3]
    want[:, 6:10] = T[:, 3:7]
    want[:, 10:14] = T[:, 7:11]
    assert np.array_equal(rows[:n], want)


@pytest.mark.gpu
@pytest.mark.parametrize("L,nsteps", [(256, 4), (2048, 2)])
def test_gpu_step2_without_grazers_equals_the_spb1_step(L, nsteps):
    """Size-independent link between the two pinned paths: with G == 0 the vegetation + grazer step reduces to the SPB = 1
    Rietkerk step (f_2Dor_2Sp R/:2460-2468 with G = 0 is f_2Dor R/:2376-2382 term by term; the vegetation keeps its Philox
    streams), so the resident SPB = 2 stepper and the strict SPB = 1 lattice kernels -- different kernels, different data
    layouts -- must agree bit for bit on P, W, O from the same seed."""
    par = RUN2["mixed_par"].astype(np.float64)
    veg, graz, cons = product_params(L, par)
    R = int(float(par[20]) / float(par[22]))
    central = np.ascontiguousarray(GOLD["central_R%d" % R], dtype=np.int32)
    frame = O.synthetic_rietkerk_frame(L, seed=77)                      # [3, L, L] = P, W, O with an extinction fringe
    f4 = np.zeros((4, L * L))
    f4[0], f4[2], f4[3] = frame[0].ravel(), frame[1].ravel(), frame[2].ravel()
    ops = tm.ConsumerOps(L)
    ops.step2_setup(veg, graz, cons, central, 1.0)
    ops.step2_upload(f4, None)
    ops.step2_run(99, trial=5, step0=0, nsteps=nsteps)
    got = ops.step2_download()[0]
    ops.close()
    lat = tm.Lattice(veg, 1, device=0)
    lat.set_strict(True)
    lat.seed(99, trial0=5)
    lat.upload(frame[None])
    lat.step(nsteps)
    want = lat.download()[0].reshape(3, L * L)
    lat.close()
    assert not got[1].any()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[1]) and np.array_equal(got[3], want[2])


@pytest.mark.parametrize("L", [20, 50])
def test_emulated_fused_rk4_equals_the_four_stage_launches(emu, L):
    """cons_rk4_fused_kernel (all four stages + clamp in one tile-fused launch, the default) against the four
    cons_rk4_stage_kernel launches: identical bits (L = 20 is smaller than a tile's cone: the periodic wrap maps several
    cone cells onto one lattice cell; L = 50 has partial edge tiles)."""
    par, pv, pg, pc, central, fr, f = setup(L, 70 + L)
    gamma = np.zeros(L * L)
    a = emu_step2(emu, L, par, central, fr, f, gamma, nsteps=2, seed=3, rk4_stages=0)
    b = emu_step2(emu, L, par, central, fr, f, gamma, nsteps=2, seed=3, rk4_stages=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.gpu
@pytest.mark.parametrize("L,nsteps", [(20, 3), (333, 3), (1024, 2)])
def test_gpu_fused_rk4_equals_the_four_stage_launches(L, nsteps):
    par, pv, pg, pc, central, fr, f = setup(L, 70 + L)
    ops = tm.ConsumerOps(L)
    ops.step2_setup(*product_params(L, par), central, fr)
    outs = []
    for staged in (0, 1):
        ops.step2_set_rk4_stages(staged)
        ops.step2_upload(f, None)
        ops.step2_run(3, nsteps=nsteps)
        outs.append(ops.step2_download())
        assert ops.last_ms()[1] == (9 if staged else 6) * nsteps
    ops.close()
    for k in range(3):
        assert np.array_equal(outs[0][k], outs[1][k])


@pytest.mark.gpu
def test_gpu_spb2_sweep_writes_the_reference_csv(tmp_path):
    """tm_b200::spb2::first_order_critical_exp_delta_stochastic_MultiSp (R/:4379 compiled with -DSPB=2): one CSV in the
    reference's place with the 2Sp header and 16-digit rows of 14 columns, a-values in order, one row per time bin."""
    L, r, div, t_max = 20, 2, 2, 2.0
    par = RUN2["mixed_par"].astype(np.float64)
    inits = np.stack([setup(L, 90 + j)[6] for j in range(r)])
    h = host_lib()
    h.tmh_riet2_sweep.argtypes = [C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, dp, C.c_int]
    h.tmh_set_options(0, 7, 0, str(tmp_path).encode())
    assert h.tmh_riet2_sweep(L, par.ctypes.data_as(dp), div, t_max, 0.05, 0.06, r, inits.ctypes.data_as(dp), r) == 0
    d = tmp_path / "Data" / "Rietkerk" / "Stochastic" / "2Sp"
    files = list(d.glob("1stCC_Rietkerk_*_P_c_G_20_T_2_dt_0.1_D1_*_a1_0.05_a2_0.06_dx_0.1_R_2.csv"))
    assert len(files) == 1, list(d.iterdir())
    lines = files[0].read_text().splitlines()
    assert "<<G(x; t)>_x>_r" in lines[0] and lines[1].strip() == ""
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[2:]])
    per_a = len(rows) // div
    assert rows.shape[1] == 14 and per_a * div == len(rows) and per_a >= 2
    assert np.all(rows[:per_a, 0] == 0.05) and np.all(rows[per_a:, 0] == 0.06) and np.all(rows[:, 1] == r) and np.all(rows[:, 2] == L)
    assert rows[0, 3] == 0.0 and np.all(np.diff(rows[:per_a, 3]) > 0) and np.isfinite(rows).all()
    assert np.all(rows[:, 8] <= r) and np.all(rows[:, 9] <= L * L)
