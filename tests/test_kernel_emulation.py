"""The product's kernel SOURCE (csrc/tm_kernels.cuh) run on host threads (tests/emu) against the
oracle: marching pipeline, halo recomputation, strip/CTA decomposition, statistics -- without a GPU.
STRICT arithmetic must match the oracle bit for bit whatever the launch geometry."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def emu_step(emu, p, fields, nsteps, strict=1, grid=0, n_sm=148, seed=1, trial0=0, step0=0, replay=None, stride=0):
    f = np.ascontiguousarray(fields, dtype=np.float64).copy()
    nt = f.shape[0]
    stats = np.zeros((nt, 8)); cnt = (C.c_ulonglong * 2)()
    rc = emu.emu_step(C.byref(p), nt, f.ctypes.data_as(dp), nsteps, strict, grid, n_sm, C.c_ulonglong(seed), trial0, step0,
                      replay.ctypes.data_as(dp) if replay is not None else None, stride, stats.ctypes.data_as(dp), cnt)
    assert rc == 0
    return f, stats, (cnt[0], cnt[1])


def oracle_steps(p, fields, nsteps, seed, trial, step0=0):
    op = O.to_oparams(p); f = np.ascontiguousarray(fields, dtype=np.float64).copy()
    for s in range(nsteps):
        rng = O.make_rng(O.SRC_PHILOX, seed=seed, trial=trial, step=step0 + s)
        O.step(op, rng if p.model != tm.MODEL_RIETKERK_DET else None, f.reshape(f.shape[0], -1))
    return f


@pytest.mark.parametrize("L,grid", [(24, 1), (40, 3), (170, 4)])
def test_rietkerk_strict_any_geometry(emu, L, grid):
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=11 + t) for t in range(2)])
    got, stats, cnt = emu_step(emu, p, frames, 2, strict=1, grid=grid, seed=2024, trial0=5)
    assert cnt == (0, 0)
    for t in range(2):
        want = oracle_steps(p, frames[t], 2, 2024, 5 + t)
        assert np.array_equal(got[t], want)
        for f in range(3):
            assert stats[t, 3 + f] == O.occupied(got[t, f])
            assert abs(stats[t, f] - O.mean(got[t, f])) <= 1e-12 * abs(O.mean(got[t, f]))


def test_split_pair_equals_fused(emu):
    """The production pair (riet_dornic_kernel -> rho* scratch -> RK4 kernel) is bit-identical to the fused kernel."""
    L = 72
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=21 + t) for t in range(2)])
    emu.emu_set_split(1)
    try:
        got, stats, cnt = emu_step(emu, p, frames, 2, strict=1, grid=3, seed=77, trial0=2)
    finally:
        emu.emu_set_split(0)
    assert cnt == (0, 0)
    for t in range(2):
        assert np.array_equal(got[t], oracle_steps(p, frames[t], 2, 77, 2 + t))


def test_result_independent_of_cta_count(emu):
    L = 48
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=2)[None]
    a, _, _ = emu_step(emu, p, frame, 3, strict=0, grid=1, seed=9)
    b, _, _ = emu_step(emu, p, frame, 3, strict=0, grid=5, seed=9)
    assert np.array_equal(a, b)


def test_rietkerk_fast_within_tolerance(emu):
    L = 40
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=3)
    got, _, _ = emu_step(emu, p, frame[None], 1, strict=0, grid=2, seed=5)
    want = oracle_steps(p, frame, 1, 5, 0)
    for f in range(3):
        scale = np.maximum(np.abs(want[f]), 1e-12 * np.abs(want[f]).max())
        assert np.max(np.abs(got[0, f] - want[f]) / scale) <= 1e-12


def test_noise_off_strict(emu):
    L = 36
    p = tm.rietkerk_params(L, model=tm.MODEL_RIETKERK_DET)
    frame = O.synthetic_rietkerk_frame(L, seed=4)
    got, _, _ = emu_step(emu, p, frame[None], 6, strict=1, grid=2)
    assert np.array_equal(got[0], oracle_steps(p, frame, 6, 0, 0))


@pytest.mark.parametrize("cubic", [False, True])
def test_dp_models(emu, cubic):
    L = 32
    if cubic:
        p = tm.dp_params(L, a=-0.72, b=-2.0, c=1.0, D=6.25, sigma=1.0, dt=0.1, dx=2.5)
        rho = np.full((L, L), 0.25)
    else:
        p = tm.dp_params(L, a=1.0548, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0)
        rho = np.ones((L, L))
    rho[:, :8] = 0.0
    got, stats, cnt = emu_step(emu, p, rho[None, None], 3, grid=2, seed=77, trial0=3)
    want = oracle_steps(p, rho[None], 3, 77, 3)
    assert np.array_equal(got[0], want)
    assert stats[0, 3] == O.occupied(want[0])


def test_replay_mode(emu):
    L, stride = 24, 96
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=8)
    u = np.random.RandomState(5).random_sample(L * L * stride)
    got, _, cnt = emu_step(emu, p, frame[None], 1, strict=1, grid=2, replay=u, stride=stride)
    assert cnt[0] == 0
    org = O.make_rng(O.SRC_REPLAY, u=u, stride=stride)
    want = frame.copy()
    O.step(O.to_oparams(p), org, want.reshape(3, -1))
    assert np.array_equal(got[0], want)


# ---- the production pair: riet_dornic_kernel + the two-column RK4 kernel (csrc/tm_rk4.cuh) -------------------------
def emu_pair(emu, *args, **kw):
    emu.emu_set_split(2)
    try:
        return emu_step(emu, *args, **kw)
    finally:
        emu.emu_set_split(0)


@pytest.mark.parametrize("L,grid,nt", [(24, 1, 2), (33, 2, 1), (72, 3, 2), (251, 5, 1), (300, 7, 1), (520, 6, 1)])
def test_rk4_pair_strict_any_geometry(emu, L, grid, nt):
    """One strip (L <= 248), odd L (scalar commit path, odd pitch), two and three strips, several row segments per strip."""
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=31 + t) for t in range(nt)])
    nsteps = 2 if L < 300 else 1
    got, stats, cnt = emu_pair(emu, p, frames, nsteps, strict=1, grid=grid, seed=909, trial0=4)
    assert cnt == (0, 0)
    for t in range(nt):
        want = oracle_steps(p, frames[t], nsteps, 909, 4 + t)
        assert np.array_equal(got[t], want), (L, t, np.abs(got[t] - want).max())
        for f in range(3):
            assert stats[t, 3 + f] == O.occupied(got[t, f])


def test_rk4_pair_fast_equals_fused_fast_and_tolerance(emu):
    """FAST arithmetic: the pair and the single fused kernel share rhs<> and the update expressions, so they agree bit for
    bit; both are within 1e-12 of the oracle."""
    L = 64
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=3)
    a, _, cnt = emu_pair(emu, p, frame[None], 2, strict=0, grid=3, seed=5)
    b, _, _ = emu_step(emu, p, frame[None], 2, strict=0, grid=2, seed=5)
    assert cnt == (0, 0) and np.array_equal(a, b)
    one, _, _ = emu_pair(emu, p, frame[None], 1, strict=0, grid=2, seed=5)
    want = oracle_steps(p, frame, 1, 5, 0)
    for f in range(3):
        scale = np.maximum(np.abs(want[f]), 1e-12 * np.abs(want[f]).max())
        assert np.max(np.abs(one[0, f] - want[f]) / scale) <= 1e-12


def test_rk4_pair_independent_of_cta_count(emu):
    L = 96
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=2)[None]
    a, _, _ = emu_pair(emu, p, frame, 3, strict=0, grid=1, seed=9)
    b, _, _ = emu_pair(emu, p, frame, 3, strict=0, grid=11, seed=9)
    assert np.array_equal(a, b)
