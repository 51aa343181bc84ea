"""Consumer-side operators of the SPB = 2 model (first part): the oracle restatement against the reference's own
outputs (tests/golden/consumer_ops.npz, made by tests/golden/make_consumer_golden.py from oracle/_ref/libtm_ref_riet2.so)."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consumer_ops.npz"))


@pytest.mark.parametrize("name", ["gamma_full", "gamma_sparse", "gamma_small", "gamma_dead"])
def test_gamma_perception_matches_reference(name):
    L, R, fr, nveg = GOLD[name + "_cfg"]
    L, R = int(L), int(R)
    central = GOLD["central_R%d" % R]
    new_length, eff = O.gamma_window(len(central), float(fr), float(nveg))
    avg = GOLD[name + "_avg"]
    got = O.gamma_perception(L, GOLD[name + "_veg"], central[:new_length], eff, avg[0], avg[1], GOLD[name + "_g0"])
    want = GOLD[name + "_out"]
    assert np.array_equal(got, want), np.abs(got - want).max()     # same additions in the same order: bit for bit
    assert got.min() >= 0.0 and got.max() <= 1.0


@pytest.mark.parametrize("name", ["rk4_a", "rk4_b"])
def test_rk4_2sp_matches_reference(name):
    L, steps = [int(x) for x in GOLD[name + "_cfg"]]
    p = O.cons_params(L, GOLD["rk4_par"])
    got = O.rk4_2sp(p, GOLD[name + "_in"], steps)
    want = GOLD[name + "_out"]
    # the reference build contracts a*b+c into FMAs, the oracle does not: <= 1e-12 relative (north_star's field tolerance)
    scale = np.abs(want).max(axis=1, keepdims=True)
    assert (np.abs(got - want) / scale).max() <= 1e-12
    assert np.array_equal(got[1] == 0.0, want[1] == 0.0)           # empty grazer sites stay empty (no diffusion in the RK4 part)


RUN2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_2sp.npz"))


@pytest.mark.parametrize("name", ["mixed", "no_grazers"])
def test_spb2_runner_frames_match_reference(name):
    """rietkerk_Dornic_2D_MultiSp (R/:3397) compiled with -DSPB=2 under sequential replay: vegetation Dornic, perception
    field, grazer Dornic with upwind advection, four-field RK4, clamp -- the oracle consumes the same uniforms in the same
    order and reproduces the runner's full-precision frames.  This is what pins the grazer's Dornic loop (inlined in the
    runner, R/:3908-4114) to the reference."""
    L = int(RUN2[name + "_L"])
    seed, n_u = [int(x) for x in RUN2[name + "_seed"]]
    pv, pg, pc, R, fr = O.runner2_params(L, RUN2[name + "_par"])
    central = GOLD["central_R%d" % R]
    u = np.random.RandomState(seed).random_sample(n_u)
    rng = O.make_rng(O.SRC_SEQ, u=u, pos=1)        # the runner draws thrID before the time loop (R/:3634)
    f = RUN2[name + "_init"].astype(np.float64).copy()
    gamma = np.zeros(L * L)
    frames, times = RUN2[name + "_frames"], RUN2[name + "_times"]
    dt = pv.dt
    t, fi = 0.0, 0
    while fi < len(times):
        if abs(t - times[fi]) < 1e-9:
            ref = frames[fi]
            for s_ in (0, 1):
                assert np.array_equal(f[s_] > 0, ref[s_] > 0), "occupancy of field %d at t=%g" % (s_, t)
            err = np.abs(f - ref) / np.maximum(np.abs(ref), 1e-300)
            err[(ref == 0) & (f == 0)] = 0
            assert err.max() <= 1e-12, "frame %d (t=%g): %.3e" % (fi, t, err.max())
            fi += 1
        O.step_2sp(pv, pg, pc, central, fr, rng, f, gamma, trig_libm=True)
        t += dt
    t_max = dt * len(times)                       # t_meas.back() of the golden run (make_runner2_golden.py)
    while t < t_max + dt:                         # the runner's loop condition, R/:3648
        O.step_2sp(pv, pg, pc, central, fr, rng, f, gamma, trig_libm=True)
        t += dt
    assert rng.overflow == 0
    # every uniform the runner drew in its time loop is accounted for; its last one names the output file (jID, R/:4343)
    assert rng.pos == int(RUN2[name + "_consumed"]) - 1
