"""SPB = 3 (vegetation + grazer + predator; VERDICT r1 missing #3): the oracle's restatement of calc_gamma_3Sp_NonRefugia (R/:3153),
RK4_Integrate_Stochastic_MultiSp over f_2Dor_3Sp (R/:2607-2720) and of one time step of rietkerk_Dornic_2D_MultiSp compiled with
-DSPB=3 against outputs of the UNMODIFIED reference (tests/golden/runner_3sp.npz, made by tests/golden/make_runner3_golden.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O

G3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_3sp.npz"))
C = O.C


@pytest.mark.parametrize("tag", ["all", "nopred"])
def test_gamma_3sp_matches_reference(tag):
    L, fr, nveg = int(G3["gamma_L"]), G3["gamma_fr"], float(G3["gamma_nveg"])
    f, avg = G3["gamma_fields"], G3["gamma_" + tag + "_avg"]
    R = int(round(max(G3["mixed_par"][27], G3["mixed_par"][28]) / G3["mixed_par"][30] - 0.5))
    cen = np.ascontiguousarray(G3["central_R%d" % R], dtype=np.int32)
    new_length = max(1, int(nveg * nveg * len(cen)))
    g = np.full((2, L * L), 7.0)                  # garbage on entry: the reference assigns (R/:3239, 3248)
    O.oracle().tmo_gamma_3sp(L, O.ptr(f[0]), O.ptr(f[1]), O.ptr(f[2]), cen.ctypes.data_as(C.POINTER(C.c_int32)), new_length,
                             float(fr[0]), float(fr[1]), O.ptr(np.ascontiguousarray(avg)), O.ptr(g[0]), O.ptr(g[1]))
    want = G3["gamma_" + tag]
    # the predator-discounted sum veg * (1 - pred / <pred>) (R/:3232) is a multiply-add chain: the reference object is built with
    # -march=x86-64-v3, whose contractions differ from the oracle's -ffp-contract=off by an ulp in some cells
    assert np.abs(g - want).max() <= 1e-14
    if tag == "nopred":
        assert np.array_equal(g, want)
    assert g.max() <= 1.0 and g.min() >= 0.0 and g[0].std() > 0 and (g[1].std() > 0 if tag == "all" else np.all(g[1] == 0.0))


def test_rk4_3sp_matches_reference():
    L = int(G3["gamma_L"])
    dt, dx, a, c, gmax, alpha, rW, W0, K1, K2, D3, D4, A01, H01, E1, A12, H12, E2 = [float(x) for x in G3["rk4_par"]]
    pc = O.OCons3Params(L, 0, dt, a, c, gmax, alpha, rW, W0, K1, K2, D3 / (dx * dx), D4 / (dx * dx), A01, H01, E1, A12, H12, E2)
    f = G3["rk4_in"].copy()
    for _ in range(3):
        assert O.oracle().tmo_rk4_3sp(C.byref(pc), O.ptr(f)) == 0
    want = G3["rk4_out3"]
    assert (np.abs(f - want) / np.abs(want).max(axis=1, keepdims=True)).max() <= 1e-13
    assert np.array_equal(f[1] == 0, want[1] == 0) and np.array_equal(f[2] == 0, want[2] == 0)


@pytest.mark.parametrize("name", ["mixed", "no_predators", "no_grazers"])
def test_spb3_runner_frames_match_reference(name):
    """vegetation Dornic, the two perception fields, grazer and predator Dornic with upwind advection, five-field RK4, clamp: the oracle
    consumes the same uniforms in the same order and reproduces the runner's full-precision frames."""
    L = int(G3[name + "_L"])
    seed, n_u = [int(x) for x in G3[name + "_seed"]]
    pv, pg1, pg2, pc, R, fr1, fr2 = O.runner3_params(L, G3[name + "_par"])
    central = G3["central_R%d" % R]
    u = np.random.RandomState(seed).random_sample(n_u)
    rng = O.make_rng(O.SRC_SEQ, u=u, pos=1)        # the runner draws thrID before the time loop (R/:3634)
    f = G3[name + "_init"].astype(np.float64).copy()
    gamma = np.zeros((2, L * L))
    frames, times = G3[name + "_frames"], G3[name + "_times"]
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
        O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, rng, f, gamma, trig_libm=True)
        t += dt
    assert fi == len(times) and rng.overflow == 0
    assert rng.pos == int(G3[name + "_consumed"]) - 1
