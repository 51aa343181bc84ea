"""Size-independent properties at BASELINE's full lattice sizes (L = 4096, and L = 1024 x many trials), where the
CPU oracle is too slow to be the checker."""
import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

pytestmark = pytest.mark.gpu


def test_L4096_seeded_run_is_reproducible_and_decomposition_invariant():
    """Same seed => same bits, whether the step is one fused launch or the Dornic + RK4 pair (different thread
    mappings, tilings and halo recomputation)."""
    L = 4096
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=77, fringe=True)
    outs = []
    for fused in (True, False, True):
        lat = tm.Lattice(p, 1)
        lat.set_fused(fused); lat.seed(2026)
        lat.upload(frame[None]); lat.step(3)
        outs.append(lat.download()[0])
        st = lat.measure()[0]
        assert st.nonfinite == 0
        lat.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    # reductions of the device agree with a host sum of the downloaded state
    for f in range(3):
        assert st.active[f] == int((outs[2][f] > 0).sum())
        assert abs(st.mean[f] - outs[2][f].mean()) <= 1e-12 * abs(outs[2][f].mean())
    assert np.all(outs[0][0][:, : L // 8] >= 0)


def test_L4096_noise_off_translation_invariance():
    """The noise-off step has no position-dependent input: rolling the frame rolls the result, bit for bit
    (periodic boundaries, strip/halo logic, ghost columns)."""
    L = 4096
    p = tm.rietkerk_params(L, model=tm.MODEL_RIETKERK_DET)
    rng = np.random.default_rng(5)
    frame = O.synthetic_rietkerk_frame(L, seed=9, fringe=True)
    frame[1] *= 1 + 0.1 * rng.random((L, L)); frame[2] *= 1 + 0.1 * rng.random((L, L))
    lat = tm.Lattice(p, 1)
    lat.upload(frame[None]); lat.step(4)
    base = lat.download()[0]
    shifted = np.roll(frame, (1237, -311), axis=(1, 2))
    lat.upload(np.ascontiguousarray(shifted)[None]); lat.step(4)
    got = lat.download()[0]
    lat.close()
    assert np.array_equal(got, np.roll(base, (1237, -311), axis=(1, 2)))


def test_L4096_uniform_state_follows_the_scalar_ode():
    """A spatially uniform noise-off state stays uniform (all Laplacians vanish) and follows the 3-variable RK4."""
    L = 4096
    p = tm.rietkerk_params(L, model=tm.MODEL_RIETKERK_DET)
    frame = np.empty((3, L, L)); frame[0] = 12.5; frame[1] = 0.31; frame[2] = 5.0
    lat = tm.Lattice(p, 1)
    lat.set_strict(True)
    lat.upload(frame[None]); lat.step(10)
    got = lat.download()[0]
    lat.close()
    small = np.empty((3, 8, 8)); small[0] = 12.5; small[1] = 0.31; small[2] = 5.0
    op = O.to_oparams(tm.rietkerk_params(8, model=tm.MODEL_RIETKERK_DET))
    for _ in range(10):
        O.step(op, None, small.reshape(3, -1))
    for f in range(3):
        assert np.all(got[f] == got[f][0, 0])
        assert got[f][0, 0] == small[f][0, 0]


def test_L1024_ensemble_trials_are_independent_of_batching():
    """config 3 shape: a trial's trajectory depends on (seed, trial id) only, not on how many trials share the launch."""
    L = 1024
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=300 + t, fringe=False) for t in range(6)])
    big = tm.Lattice(p, 6); big.seed(8, trial0=10); big.upload(frames); big.step(3)
    all6 = big.download(); big.close()
    one = tm.Lattice(p, 1); one.seed(8, trial0=14); one.upload(frames[4:5]); one.step(3)
    assert np.array_equal(one.download()[0], all6[4])
    one.close()


def test_dp_L2048_sweep_extinction_is_absorbing_and_monotone_in_a():
    """DP a-sweep (config 4 shape): the absorbing state is absorbing; larger a keeps more active sites."""
    L = 2048
    p = tm.dp_params(L, a=0.5, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0)
    avals = [0.2, 0.9, 1.0548, 1.3]
    lat = tm.Lattice(p, len(avals) + 1)
    for i, a in enumerate(avals):
        lat.set_trial_a(i, a)
    rho0 = np.ones((len(avals) + 1, 1, L, L)); rho0[-1] = 0.0      # last trial starts extinct
    lat.seed(1); lat.upload(rho0); lat.step(30)
    st = lat.measure()
    assert st[len(avals)].active[0] == 0 and st[len(avals)].mean[0] == 0.0
    act = [st[i].active[0] for i in range(len(avals))]
    assert act[0] < act[1] < act[2] < act[3] and all(s.nonfinite == 0 and s.negative == 0 for s in st)
    lat.close()
