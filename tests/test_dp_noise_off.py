"""The noise-off DP path (second half): percNoDiff_2D, Percolation_FD/stochasticSPDP.cpp:549-662 -- classical RK4
over fP_2D (:516-529) with the fourth-order nine-point-cross Laplacian of determine_neighbours_R2 (:394-414).

CPU: the oracle restatement against the rows of the UNMODIFIED reference runner (tests/golden/dp_noise_off.npz, made by
tests/golden/make_dp_noise_off_golden.py; <= 1e-12: the reference binary contracts a*b+c into FMAs, the oracle does not),
and the product's stage kernel SOURCE on host threads against the oracle (bit for bit).
GPU (-m gpu): the C ABI tm_dp_noise_off_host against the oracle (bit for bit), the host runner with the reference's
signature against the reference's rows, and size-independent properties at L = 2048."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "dp_noise_off.npz"))
CASES = ["growth", "decay_patchy"]
dp = C.POINTER(C.c_double)


def case(name):
    L, a, b, D, dt, dx = [float(x) for x in GOLD[name + "_par"]]
    return int(L), a, b, D, dt, dx, GOLD[name + "_rho0"], GOLD[name + "_tstop"], GOLD[name + "_rows"]


def replay_schedule(ts, dt, t_max):
    """Steps after which percNoDiff_2D logs (accumulated t += dt, window P/:638, s cycling P/:646) -> [(nsteps so far, t)]."""
    t, s, k, out = 0.0, 0, 0, []
    while t < t_max + dt:
        k += 1
        t += dt
        if t > ts[s] - dt / 2.0 and t <= ts[s] + dt / 2.0:
            out.append((k, t))
            s = (s + 1) % len(ts)
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_reference_runner(name):
    L, a, b, D, dt, dx, rho0, ts, rows = case(name)
    sched = replay_schedule(ts, dt, float(ts[-1]))
    assert len(sched) * L * L == len(rows)
    rho, done = rho0.copy(), 0
    for w, (k, t) in enumerate(sched):
        rho = O.dp_noise_off(L, a, b, D, dt, dx, rho, k - done)
        done = k
        blk = rows[w * L * L:(w + 1) * L * L]
        assert np.all(blk[:, 0] == a) and np.all(blk[:, 1] == 0) and np.array_equal(blk[:, 3], np.arange(L * L))
        assert np.all(blk[:, 2] == t)                      # the same accumulated time, bit for bit
        ref = blk[:, 4]
        assert np.abs(rho - ref).max() <= 1e-12 * np.abs(ref).max(), "window %d: %.3e" % (w, np.abs(rho - ref).max())


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    e = C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))
    e.emu_dp_noise_off.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, C.c_int]
    return e


@pytest.mark.parametrize("name", CASES)
def test_emulated_stage_kernel_equals_the_oracle(emu, name):
    L, a, b, D, dt, dx, rho0, ts, rows = case(name)
    got = rho0.copy()
    emu.emu_dp_noise_off(L, a, b, D, dt, dx, got.ctypes.data_as(dp), 7)
    assert np.array_equal(got, O.dp_noise_off(L, a, b, D, dt, dx, rho0, 7))


def test_emulated_kernel_on_tiny_and_odd_lattices(emu):
    """L < 5: the radius-2 neighbours wrap onto each other (and onto the cell itself at L = 1, 2); L = 131: a ragged last tile."""
    for L in (1, 2, 3, 4, 131):
        rho0 = np.random.default_rng(L).random(L * L)
        got = rho0.copy()
        emu.emu_dp_noise_off(L, 0.5, 1.0, 0.3, 0.05, 1.0, got.ctypes.data_as(dp), 2)
        assert np.array_equal(got, O.dp_noise_off(L, 0.5, 1.0, 0.3, 0.05, 1.0, rho0, 2))


def test_oracle_homogeneous_field_follows_the_scalar_rk4():
    """No gradients: the Laplacian weights cancel exactly (4 - 64 + 60 = 0) and every cell follows the scalar RK4 of a r - b r^2."""
    L, a, b, dt = 16, 0.8, 1.3, 0.1
    rho = O.dp_noise_off(L, a, b, 0.7, dt, 1.0, np.full(L * L, 0.2), 10)
    r = 0.2
    f = lambda x: a * x - b * x * x
    for _ in range(10):
        k1 = f(r); k2 = f(r + dt / 2 * k1); k3 = f(r + dt / 2 * k2); k4 = f(r + dt * k3)
        r += dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert np.abs(rho - r).max() <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_equals_the_oracle(name):
    L, a, b, D, dt, dx, rho0, ts, rows = case(name)
    got, ms = tm.dp_noise_off(L, a, b, D, dt, dx, rho0, 25)
    assert np.array_equal(got, O.dp_noise_off(L, a, b, D, dt, dx, rho0, 25)) and ms > 0
    for Lx in (1, 3, 131, 300):
        r0 = np.random.default_rng(Lx).random(Lx * Lx)
        assert np.array_equal(tm.dp_noise_off(Lx, 0.5, 1.0, 0.3, 0.05, 1.0, r0, 3)[0], O.dp_noise_off(Lx, 0.5, 1.0, 0.3, 0.05, 1.0, r0, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_host_runner_rows_match_the_reference(name):
    """tm_b200::percNoDiff_2D (the reference's argument list) against the rows of the reference's own runner."""
    L, a, b, D, dt, dx, rho0, ts, rows = case(name)
    h = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_dp_noise_off.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, dp, C.c_int, C.c_double, dp, C.c_int]
    out = np.zeros((len(rows) + 8, 5))
    n = h.tmh_dp_noise_off(L, a, b, D, dt, dx, rho0.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts), float(ts[-1]),
                           out.ctypes.data_as(dp), len(out))
    assert n == len(rows)
    assert np.array_equal(out[:n, :4], rows[:, :4])                       # a, replicate, t, cell: exact
    assert np.abs(out[:n, 4] - rows[:, 4]).max() <= 1e-12 * np.abs(rows[:, 4]).max()


@pytest.mark.gpu
def test_gpu_fullsize_properties():
    """L = 2048: translation covariance under np.roll (bit for bit: the stencil is the same everywhere), a homogeneous field
    stays homogeneous and follows the scalar RK4, run-to-run reproducibility."""
    L, a, b, D, dt, dx = 2048, 1.1, 1.0, 0.25, 0.05, 1.0
    rho0 = np.random.default_rng(3).random((L, L))
    got, ms = tm.dp_noise_off(L, a, b, D, dt, dx, rho0, 4)
    again, _ = tm.dp_noise_off(L, a, b, D, dt, dx, rho0, 4)
    assert np.array_equal(got, again)
    rolled, _ = tm.dp_noise_off(L, a, b, D, dt, dx, np.roll(rho0, (5, -9), axis=(0, 1)), 4)
    assert np.array_equal(rolled.reshape(L, L), np.roll(got.reshape(L, L), (5, -9), axis=(0, 1)))
    flat, _ = tm.dp_noise_off(L, a, b, D, dt, dx, np.full(L * L, 0.3), 4)
    assert np.all(flat == flat[0]) and flat[0] == O.dp_noise_off(4, a, b, D, dt, dx, np.full(16, 0.3), 4)[0]
    assert np.isfinite(got).all()
