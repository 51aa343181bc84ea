"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fields <= 1e-12 relative in FP64; integer outputs
(active-site counts, sampler decisions) bit-exact.  In STRICT mode the kernels keep the
reference's operation order and must match the oracle bit for bit.
"""
import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def rel_err(got, want, floor=1e-300):
    """max |got-want| / max(|want|, scale*1e-4): relative to the field scale for tiny cells."""
    got, want = np.asarray(got), np.asarray(want)
    scale = np.maximum(np.abs(want), floor)
    return float(np.max(np.abs(got - want) / scale))


def oracle_steps(p, fields, nsteps, seed, trial, step0=0):
    """fields [n_fields][L][L] -> oracle state after nsteps with the Philox stream."""
    op = O.to_oparams(p)
    f = np.ascontiguousarray(fields, dtype=np.float64).copy()
    for s in range(nsteps):
        rng = O.make_rng(O.SRC_PHILOX, seed=seed, trial=trial, step=step0 + s)
        O.step(op, rng if p.model != tm.MODEL_RIETKERK_DET else None, f.reshape(f.shape[0], -1))
    return f


@pytest.mark.parametrize("L", [32, 64, 100, 250, 400])
def test_rietkerk_strict_bit_exact(L):
    p = tm.rietkerk_params(L)
    n_trials = 2
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=11 + t) for t in range(n_trials)])
    lat = tm.Lattice(p, n_trials)
    lat.set_strict(True)
    lat.seed(2024, trial0=5)
    lat.upload(frames)
    nsteps = 3
    lat.step(nsteps)
    got = lat.download()
    assert lat.sampler_overflow() == 0
    for t in range(n_trials):
        want = oracle_steps(p, frames[t], nsteps, 2024, 5 + t)
        assert np.array_equal(got[t], want), "trial %d: max rel %.3e" % (t, rel_err(got[t], want))
    st = lat.measure()
    for t in range(n_trials):
        for f in range(3):
            assert st[t].active[f] == int(O.occupied(got[t, f]))
            assert abs(st[t].mean[f] - O.mean(got[t, f])) <= RTOL * abs(O.mean(got[t, f]))
        assert st[t].nonfinite == 0
    lat.close()


def test_split_pair_equals_fused_kernel():
    """tm_set_fused(0): Dornic sweep + RK4 kernel pair, bit-identical to the fused kernel (strict and fast)."""
    L = 200
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=60 + t) for t in range(2)])
    out = {}
    for strict in (True, False):
        for fused in (True, False):
            lat = tm.Lattice(p, 2)
            lat.set_strict(strict); lat.set_fused(fused); lat.seed(5150, trial0=3)
            lat.upload(frames); lat.step(4)
            out[(strict, fused)] = lat.download()
            assert lat.sampler_overflow() == 0
            lat.close()
        assert np.array_equal(out[(strict, True)], out[(strict, False)])
    want = oracle_steps(p, frames[1], 4, 5150, 4)
    assert np.array_equal(out[(True, False)][1], want)


@pytest.mark.parametrize("L,n_trials", [(8, 3), (9, 2), (17, 5), (33, 300)])
def test_rietkerk_small_and_odd_lattices(L, n_trials):
    """Edge geometries: minimum L, odd L (ragged strips, odd pitch), many more trials than SMs."""
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=100 + t) for t in range(n_trials)])
    lat = tm.Lattice(p, n_trials)
    lat.set_strict(True)
    lat.seed(11, trial0=7)
    lat.upload(frames)
    lat.step(2)
    got = lat.download()
    st = lat.measure()
    for t in sorted(set([0, 1, n_trials // 2, n_trials - 1])):
        want = oracle_steps(p, frames[t], 2, 11, 7 + t)
        assert np.array_equal(got[t], want), "trial %d" % t
        assert st[t].active[0] == int(O.occupied(want[0]))
    lat.close()


def test_extinct_lattice_stays_extinct_and_empty_inputs():
    """All-zero vegetation: the skip rule (R/:3759) draws nothing, P stays exactly 0, water relaxes."""
    L = 40
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=1)
    frame[0] = 0.0
    lat = tm.Lattice(p, 1)
    lat.seed(3)
    lat.upload(frame[None])
    lat.step(5)
    got = lat.download()[0]
    assert np.all(got[0] == 0.0)
    assert np.array_equal(got, oracle_steps(p, frame, 5, 3, 0)) or np.allclose(got, oracle_steps(p, frame, 5, 3, 0), rtol=1e-12)
    st = lat.measure()
    assert st[0].active[0] == 0 and st[0].active[1] == L * L
    lat.step(0)   # zero steps is a no-op
    lat.close()


@pytest.mark.parametrize("L", [64, 250])
def test_rietkerk_fast_per_step(L):
    """Per-step parity (north_star): each GPU step starts from the oracle's state."""
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=3)
    lat = tm.Lattice(p, 1)
    lat.seed(77)
    state = frame.copy()
    for s in range(6):
        lat.upload(state[None])
        lat.set_step_counter(s)
        lat.step(1)
        got = lat.download()[0]
        state = oracle_steps(p, state, 1, 77, 0, step0=s)
        for f in range(3):
            scale = np.maximum(np.abs(state[f]), 1e-12 * np.abs(state[f]).max())
            err = np.abs(got[f] - state[f]) / scale
            # <= 1e-12 everywhere, except that a 1-ulp difference in the Gaussian-Poisson argument
            # may cross the integer truncation `long(norm*sqrt(z)+z)` (R/:3776): a 1/g ~ 1e-8 jump,
            # expected in ~3e-8 of the cells per step.  Allow at most 3 such cells.
            assert (err > RTOL).sum() <= 3 and err.max() <= 1e-6, "step %d field %d err %.3e (%d cells)" % (
                s, f, err.max(), (err > RTOL).sum())
        assert np.array_equal(got[0] > 0, state[0] > 0)
    lat.close()


@pytest.mark.parametrize("strict", [True, False])
def test_rietkerk_noise_off(strict):
    """Noise-off path (f_2D + RK4, R/:1718-1814): 40 free-running steps."""
    L = 96
    p = tm.rietkerk_params(L, model=tm.MODEL_RIETKERK_DET)
    frame = O.synthetic_rietkerk_frame(L, seed=5)
    lat = tm.Lattice(p, 1)
    lat.set_strict(strict)
    lat.upload(frame[None])
    lat.step(40)
    got = lat.download()[0]
    want = oracle_steps(p, frame, 40, 0, 0)
    if strict:
        assert np.array_equal(got, want), rel_err(got, want)
    else:
        for f in range(3):
            scale = np.maximum(np.abs(want[f]), 1e-12 * np.abs(want[f]).max())
            assert np.max(np.abs(got[f] - want[f]) / scale) <= RTOL
    lat.close()


@pytest.mark.parametrize("cubic,L", [(False, 64), (True, 64), (False, 200), (True, 128)])
def test_dp_bit_exact(cubic, L):
    if cubic:
        p = tm.dp_params(L, a=-0.72, b=-2.0, c=1.0, D=6.25, sigma=1.0, dt=0.1, dx=2.5)
        rho0 = np.full((L, L), 0.25)
    else:
        p = tm.dp_params(L, a=1.0548, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0)
        rho0 = np.ones((L, L))
    rng = np.random.default_rng(9)
    rho0[:, : L // 4] = 0.0                       # an empty band: skip rule + front cells with mean 0
    rho0[: L // 8, :] *= rng.random((L // 8, L))  # ragged densities
    lat = tm.Lattice(p, 3)
    lat.seed(4242, trial0=1)
    lat.upload(np.stack([rho0[None]] * 3))
    lat.step(5)
    got = lat.download()
    assert lat.sampler_overflow() == 0
    for t in range(3):
        want = oracle_steps(p, rho0[None], 5, 4242, 1 + t)
        assert np.array_equal(got[t], want), "trial %d max rel %.3e" % (t, rel_err(got[t], want))
    st = lat.measure()
    for t in range(3):
        assert st[t].active[0] == int(O.occupied(got[t, 0]))
        assert abs(st[t].mean[0] - O.mean(got[t, 0])) <= RTOL * abs(O.mean(got[t, 0]))
    lat.close()


@pytest.mark.parametrize("model", ["riet", "dp", "cubic", "riet_fast"])
def test_replay_host_uniforms(model):
    """Replay mode: host-supplied uniforms, fixed per-cell stride, one step.  `riet_fast` is the same replay through the
    FAST arithmetic (the benchmarked kernels' arithmetic): north_star's replay criterion, <= 1e-12."""
    fast = model == "riet_fast"
    if fast:
        model = "riet"
    L, stride = 48, 96
    rng = np.random.default_rng(21)
    if model == "riet":
        p = tm.rietkerk_params(L)
        frame = O.synthetic_rietkerk_frame(L, seed=8)
    elif model == "dp":
        p = tm.dp_params(L, a=1.0548, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0)
        frame = rng.random((1, L, L)); frame[0, :, :10] = 0
    else:
        p = tm.dp_params(L, a=-0.72, b=-2.0, c=1.0, D=6.25, sigma=1.0, dt=0.1, dx=2.5)
        frame = rng.random((1, L, L)) * 1.5; frame[0, :, :10] = 0
    u = (rng.integers(0, 2 ** 52, L * L * stride, dtype=np.uint64).astype(np.float64)) * 2.0 ** -52
    lat = tm.Lattice(p, 1)
    lat.set_strict(not fast)
    lat.upload(frame[None])
    lat.step_replay(u, stride)
    got = lat.download()[0]
    assert lat.sampler_overflow() == 0
    org = O.make_rng(O.SRC_REPLAY, u=u, stride=stride)
    want = frame.copy()
    O.step(O.to_oparams(p), org, want.reshape(want.shape[0], -1))
    assert org.overflow == 0
    if fast:
        for f in range(3):
            scale = np.maximum(np.abs(want[f]), 1e-12 * np.abs(want[f]).max())
            err = np.abs(got[f] - want[f]) / scale
            assert err.max() <= RTOL, "field %d: %.3e" % (f, err.max())   # 2304 cells: no truncation flip expected (3e-8 per cell)
        assert np.array_equal(got[0] > 0, want[0] > 0)
    else:
        assert np.array_equal(got, want), rel_err(got, want)
    lat.close()


def test_dornic_only_matches_oracle():
    L = 64
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=2)
    lat = tm.Lattice(p, 1)
    lat.seed(5)
    lat.upload(frame[None])
    got = lat.dornic_only()[0]
    rng = O.make_rng(O.SRC_PHILOX, seed=5, trial=0, step=0)
    want = O.dornic_only(O.to_oparams(p), rng, frame[0].ravel().copy()).reshape(L, L)
    assert np.array_equal(got, want), rel_err(got, want)
    # state untouched
    assert np.array_equal(lat.download()[0], frame)
    lat.close()


def test_decomposition_invariance():
    """Row slabs + halo exchange (host-staged here) reproduce the whole-lattice result bit for bit."""
    import ctypes as C
    L, nslab = 64, 4
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=13)
    whole = tm.Lattice(p, 1)
    whole.seed(99)
    whole.upload(frame[None])
    whole.step(3)
    want = whole.download()[0]
    whole.close()
    rows = L // nslab
    slabs = [tm.Lattice(p, 1, row0=k * rows, nrows=rows) for k in range(nslab)]
    cudart = C.CDLL("libcudart.so")
    for k, s in enumerate(slabs):
        s.seed(99)
        s.upload(np.ascontiguousarray(frame[:, k * rows:(k + 1) * rows][None]))
    nbytes = slabs[0].halo_bytes()
    for step in range(3):
        tops = [s.halo_pack(0) for s in slabs]
        bots = [s.halo_pack(1) for s in slabs]
        for s in slabs:
            s.sync()
        for k, s in enumerate(slabs):
            s.halo_unpack(0, bots[(k - 1) % nslab])   # rows above me = bottom rows of the slab above
            s.halo_unpack(1, tops[(k + 1) % nslab])
        for s in slabs:
            s.sync()
        for s in slabs:
            s.step(1)
    got = np.concatenate([s.download()[0] for s in slabs], axis=1)
    assert np.array_equal(got, want)
    for s in slabs:
        s.close()


def test_slab_decomposition_over_nccl():
    """BASELINE config 5 in small: row slabs on 2+ GPUs, halo exchange over NCCL, bit-exact vs one GPU."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "run_slab_nccl.py"), "256", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "bit_exact=True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
