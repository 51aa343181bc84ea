"""The Rietkerk_FPE multi-species runner (VERDICT r1 missing #4): oracle restatement of one time step of
rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=2 (F/:4094-4462) against full-precision frames of the UNMODIFIED reference runner
under sequential replay (tests/golden/runner_fpe_2sp.npz, made by tests/golden/make_fpe_runner_golden.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O

RUN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_fpe_2sp.npz"))


@pytest.mark.parametrize("name", ["mixed", "no_grazers"])
def test_fpe_2sp_runner_frames_match_reference(name):
    """perception field -> vegetation Dornic -> four-field RK4 (grazer mortality in the reaction terms, no grazer Dornic) -> FKE
    movement of the grazer -> clamp at 1e-10: the oracle consumes the same uniforms in the same order and reproduces the runner's
    frames."""
    L = int(RUN[name + "_L"])
    seed, n_u = [int(x) for x in RUN[name + "_seed"]]
    pv, pc, M1, terms, gD, gV, R, fr = O.fpe_runner_params(L, RUN[name + "_par"], RUN[name + "_derived"])
    central = RUN["central_R%d" % R]
    u = np.random.RandomState(seed).random_sample(n_u)
    rng = O.make_rng(O.SRC_SEQ, u=u, pos=1)        # the runner draws thrID before the time loop (F/:3960)
    f = RUN[name + "_init"].astype(np.float64).copy()
    gamma = np.zeros(L * L)
    frames, times = RUN[name + "_frames"], RUN[name + "_times"]
    dt = pv.dt
    t, fi = 0.0, 0
    t_max = times[-1]                              # t_meas.back() of the golden run
    while t < t_max + dt:                          # the runner's loop condition, F/:3986
        if fi < len(times) and abs(t - times[fi]) < 1e-9:
            ref = frames[fi]
            for s_ in (0, 1):
                assert np.array_equal(f[s_] > 0, ref[s_] > 0), "occupancy of field %d at t=%g" % (s_, t)
            err = np.abs(f - ref) / np.maximum(np.abs(ref), 1e-300)
            err[(ref == 0) & (f == 0)] = 0
            assert err.max() <= 1e-12, "frame %d (t=%g): %.3e" % (fi, t, err.max())
            fi += 1
        O.step_fpe_2sp(pv, pc, M1, terms, gD, gV, central, fr, rng, f, gamma)
        t += dt
    assert fi == len(times)
    assert rng.overflow == 0
    assert rng.pos == int(RUN[name + "_consumed"]) - 1   # the runner's last draw names the output file (jID)


def test_fpe_flavours_differ_from_fd_only_where_the_reference_does():
    """gamma: a vanished consumer gets 1.0 (F/:2579) instead of 0.0 (R/:2575); RK4: the grazer's mortality term (F/:2466)"""
    L, rng = 12, np.random.default_rng(3)
    prey = rng.random(L * L)
    cen = RUN["central_R5"][:20]
    g_fd, g_fpe = np.zeros(L * L), np.zeros(L * L)
    cp = cen.ctypes.data_as(O.C.POINTER(O.C.c_int32))
    O.oracle().tmo_gamma_perception(L, O.ptr(prey), cp, 20, 0.5, 0.0, O.ptr(g_fd))
    O.oracle().tmo_gamma_perception_fpe(L, O.ptr(prey), cp, 20, 0.5, 0.0, O.ptr(g_fpe))
    assert np.all(g_fd == 0.0) and np.all(g_fpe == 1.0)
    g_fd[:] = 0; g_fpe[:] = 0
    O.oracle().tmo_gamma_perception(L, O.ptr(prey), cp, 20, 0.5, 0.3, O.ptr(g_fd))
    O.oracle().tmo_gamma_perception_fpe(L, O.ptr(prey), cp, 20, 0.5, 0.3, O.ptr(g_fpe))
    assert np.array_equal(g_fd, g_fpe) and g_fd.max() > 0
    pv, pc, M1, *_ = O.fpe_runner_params(L, RUN["mixed_par"], RUN["mixed_derived"])
    f0 = np.abs(rng.normal(1.0, 0.3, (4, L * L)))
    a, b = f0.copy(), f0.copy()
    assert O.oracle().tmo_rk4_2sp(O.C.byref(pc), O.ptr(a)) == 0
    assert O.oracle().tmo_rk4_fpe_2sp(O.C.byref(pc), 0.0, O.ptr(b)) == 0
    assert np.array_equal(a, b)                                  # M1 = 0: the FD reaction terms
    assert O.oracle().tmo_rk4_fpe_2sp(O.C.byref(pc), 0.5, O.ptr(b)) == 0
    assert np.all(b[1] < a[1])


RUN3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_fpe_3sp.npz"))
G3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_3sp.npz"))


@pytest.mark.parametrize("name", ["mixed", "no_predators"])
def test_fpe_3sp_runner_frames_match_reference(name):
    """SPB = 3: calc_gamma_vel_3Sp_NonRefugia (a direction uniform per site and species in radius order, the vegetation's included; the
    perception fields accumulated onto their old values) -> vegetation Dornic -> five-field RK4 with both mortalities -> FKE movement
    of grazer and predator -> clamp: the oracle consumes the same uniforms in the same order and reproduces the runner's frames."""
    L = int(RUN3[name + "_L"])
    seed, n_u = [int(x) for x in RUN3[name + "_seed"]]
    pv, pc, M, v, terms, gD, gV, R, fr1, fr2 = O.fpe_runner3_params(L, RUN3[name + "_par"], RUN3[name + "_derived"])
    central = G3["central_R%d" % R]
    u = np.random.RandomState(seed).random_sample(n_u)
    rng = O.make_rng(O.SRC_SEQ, u=u, pos=1)
    f = RUN3[name + "_init"].astype(np.float64).copy()
    gamma, veff = np.zeros((2, L * L)), np.zeros((2, L * L, 2))
    frames, times = RUN3[name + "_frames"], RUN3[name + "_times"]
    dt = pv.dt
    t, fi = 0.0, 0
    t_max = times[-1]
    while t < t_max + dt:
        if fi < len(times) and abs(t - times[fi]) < 1e-9:
            ref = frames[fi]
            for s_ in (0, 1, 2):
                assert np.array_equal(f[s_] > 0, ref[s_] > 0), "occupancy of field %d at t=%g" % (s_, t)
            err = np.abs(f - ref) / np.maximum(np.abs(ref), 1e-300)
            err[(ref == 0) & (f == 0)] = 0
            assert err.max() <= 1e-12, "frame %d (t=%g): %.3e" % (fi, t, err.max())
            fi += 1
        O.step_fpe_3sp(pv, pc, M, v, terms, gD, gV, central, fr1, fr2, rng, f, gamma, veff, trig_libm=True)
        t += dt
    assert fi == len(times) and rng.overflow == 0
    assert rng.pos == int(RUN3[name + "_consumed"]) - 1
