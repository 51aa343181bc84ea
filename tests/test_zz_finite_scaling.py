"""DP finite-size scaling drivers of the host layer: critical_exp_finite_scaling_collapse (P/:1424) and
critical_exp_finite_scaling_stationary (P/:1538) -- percolationDornic_2D at a_c from a full frame over a range of grid sizes, one CSV
with the first seven columns of the runner's rows.  Checked: file-name grammar, header, row counts from the reference's time-bin
rules, the L column, t = 0 rows of a full frame, and that every block equals what the runner itself returns for that grid size
(same Philox key (seed, a) per trial => the same trajectories)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)
HEADER = " a , L, t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt "


@pytest.fixture(scope="module")
def host():
    tm.lib()
    h = C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_dp_run.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_int, dp, dp, C.c_int, dp, C.c_int]
    h.tmh_dp_finite_scaling.argtypes = [C.c_int] * 4 + [C.c_double] * 7 + [C.c_int]
    return h


def runner_rows(host, L, a, b, D, sigma, dt, dx, t_meas, r):
    rows = np.zeros((4096, 9))
    tmv = np.ascontiguousarray(t_meas, dtype=np.float64)
    r0 = np.ones(L * L)
    n = host.tmh_dp_run(0, L, a, b, 0.0, D, sigma, dt, dx, r, O.ptr(r0), O.ptr(tmv), len(tmv), O.ptr(rows), 4096)
    assert n > 0, n
    return rows[:n, :8]


def test_finite_scaling_collapse_csv(host, tmp_path):
    a_c, b, D, sigma, dt, dx, r, t_max = 1.0548, 1.0, 0.25, 2 ** 0.5, 0.1, 1.0, 2, 3.0
    host.tmh_set_options(0, 11, 0, str(tmp_path).encode())
    assert host.tmh_dp_finite_scaling(0, 3, 16, 32, t_max, a_c, b, D, sigma, dt, dx, r) == 0
    path = tmp_path / "Data" / "SPDP" / "DEBUG_P_c_Finite_Scaling_Collapse_T_3_dt_0.1_a_1.0548_L1_16_L2_32_Sig_1.41_R_2.csv"
    assert path.exists(), list((tmp_path / "Data" / "SPDP").iterdir())
    lines = path.read_text().splitlines()
    assert lines[0] == HEADER
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])
    bins = np.zeros(64)
    nb = O.oracle().tmo_logarithm10_time_bins(C.c_double(t_max), C.c_double(dt), O.ptr(bins), 64)
    per_g = 11 + nb                                        # every step while t <= 1, then the log10 bins (P/:949-964)
    assert rows.shape == (3 * per_g, 7)
    assert np.all(rows[:per_g, 1] == 16) and np.all(rows[per_g:2 * per_g, 1] == 24) and np.all(rows[2 * per_g:, 1] == 32)
    assert np.all(rows[:, 0] == a_c) and np.all(rows[rows[:, 2] == 0][:, 3] == 1.0)       # init_fullframe
    for k, L in enumerate((16, 24, 32)):
        want = runner_rows(host, L, a_c, b, D, sigma, dt, dx, bins[:nb], r)
        got = rows[k * per_g:(k + 1) * per_g]
        np.testing.assert_allclose(got, want[:, :7], rtol=1e-15, atol=0)                  # setprecision(16) columns; t and a exact


def test_finite_scaling_stationary_csv(host, tmp_path, monkeypatch):
    a_c, b, D, sigma, dt, dx, r = 1.0548, 1.0, 0.25, 2 ** 0.5, 0.1, 1.0, 2
    monkeypatch.setenv("TM_FSS_WINDOW_SCALE", "0.002")     # the reference's windows are 900 .. 15000 time units long
    host.tmh_set_options(0, 12, 0, str(tmp_path).encode())
    assert host.tmh_dp_finite_scaling(1, 2, 32, 64, 50.0, a_c, b, D, sigma, dt, dx, r) == 0
    path = tmp_path / "Data" / "SPDP" / "DEBUG_P_c_Finite_Scaling_Stationary_T_50_dt_0.1_a_1.0548_L1_32_L2_64_Sig_1.41_R_2.csv"
    assert path.exists(), list((tmp_path / "Data" / "SPDP").iterdir())
    lines = path.read_text().splitlines()
    assert lines[0] == HEADER
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])
    import math
    g0 = int(math.pow(math.exp(1), math.log(32.0)))        # lnspace (P/:281-301) rebuilds its first point as e^(ln 32): 31 with glibc's pow
    assert rows.shape[1] == 7 and set(rows[:, 1]) == {float(g0), 64.0}
    for L, (t0, t1) in ((g0, (300, 900)), (64, (1000, 2500))):
        blk = rows[rows[:, 1] == L]
        tm_ = np.linspace(t0 * 0.002, t1 * 0.002, 30)
        want = runner_rows(host, L, a_c, b, D, sigma, dt, dx, tm_, r)
        assert len(blk) == len(want)
        np.testing.assert_allclose(blk, want[:, :7], rtol=1e-15, atol=0)
    # grid sizes the reference has no window for are refused, not guessed
    assert host.tmh_dp_finite_scaling(1, 2, 16, 64, 50.0, a_c, b, D, sigma, dt, dx, r) == -2


@pytest.mark.parametrize("exe,stdin,pattern,ncol", [
    ("delta_frame_test", "0.1\n3\n24\n2\n1.0524\n1.0554\n2\n", "DEBUG_P_c_Delta_DP_G_24_T_3_dt_0.1_a1_1.052_a2_1.055_Sig_1.41_R_2.csv", 8),
    ("spreading_exp_test", "0.1\n3\n25\n3\n2\n1.0548\n", "SPREADING_P_c_ThDel_DP_G_25_T_3_*", None),
    ("finite_scaling_delta", "0.1\n50\n33\n64\n2\n2\n1.0548\n", "DEBUG_P_c_Finite_Scaling_Stationary_T_50_dt_0.1_a_1.0548_L1_33_L2_64_Sig_1.41_R_2.csv", 7),
])
def test_dp_driver_binaries(tmp_path, monkeypatch, exe, stdin, pattern, ncol):
    """drivers/{delta_frame_test, spreading_exp_test, finite_scaling_delta}: the stdin protocols and constants of the reference's DP driver
    mains (Percolation_FD/<same name>.cpp) over the host layer; each writes its one CSV in the reference's place."""
    import glob
    import subprocess
    path = os.path.join(ROOT, "drivers", exe)
    assert os.path.exists(path), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7", TM_FSS_WINDOW_SCALE="0.002")
    out = subprocess.run([path], input=stdin, text=True, capture_output=True, env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    files = glob.glob(str(tmp_path / "Data" / "SPDP" / pattern))
    assert len(files) == 1, os.listdir(tmp_path / "Data" / "SPDP")
    rows = np.loadtxt(files[0], delimiter=",", skiprows=1)
    assert rows.ndim == 2 and len(rows) > 4 and np.isfinite(rows).all()
    if ncol:
        assert rows.shape[1] == ncol
    bad = subprocess.run([path], input="0.1\nnot-a-number\n", text=True, capture_output=True, env=env)
    assert bad.returncode == 2


def test_bifurc_prob_driver_and_terminal_runner(host, tmp_path):
    """drivers/bifurc_prob -> first_order_terminal_diff (P/:2548) -> terminal_percolationDornic_2D (P/:2158): one row per initial density
    and replicate, {a, c, L, j, t, <rho(0)>, <rho(t)>, #active}; an empty start stays empty; every row equals an independent run of the
    same trial through the C ABI (cubic DP step, the runner's stream key, the reference's window rule for the measurement step)."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "bifurc_prob")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    L, a, t_event, dt, r, seed = 24, -0.9, 1.0, 0.1, 2, 7
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED=str(seed))
    out = subprocess.run([exe], input="%g\n%g\n%d\n%g\n0\n1.5\n4\n%d\n" % (dt, t_event, L, a, r), text=True, capture_output=True, env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    path = tmp_path / "Data" / "SPDP" / "1stOrder_Rep_DP_G_24_T_1_dt_0.1_a_-0.9_c_1_p1_0_p2_1.5_dx_2.5_Sig_1_R_2.csv"
    assert path.exists(), os.listdir(tmp_path / "Data" / "SPDP")
    lines = path.read_text().splitlines()
    assert lines[0] == "  a , c , L, r, t ,  <<Rho(x; 0)>_x>, <<Rho(x; t)>_x>, # Active Sites "
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])
    assert rows.shape == (4 * r, 8) and np.all(rows[:, 0] == a) and np.all(rows[:, 1] == 1) and np.all(rows[:, 2] == L) and np.all(rows[:, 4] == t_event)
    assert list(rows[:, 3]) == [0, 1] * 4 and np.allclose(rows[:, 5], np.repeat([0, 0.5, 1.0, 1.5], r))
    assert np.all(rows[:r, 6:] == 0) and np.all(rows[r:, 6] > 0)
    steps, t = 0, 0.0
    while not (t != 0.0 and t_event - dt / 2.0 <= t < t_event + dt / 2.0):
        t += dt
        steps += 1
    host.tmh_stream_seed.restype = C.c_ulonglong
    host.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    for k, p0 in enumerate([0.5, 1.0, 1.5], start=1):
        lat = tm.Lattice(tm.dp_params(L, a=a, b=-2.0, D=6.25, sigma=1.0, dt=dt, dx=2.5, c=1.0), r)
        lat.seed(host.tmh_stream_seed(host.tmh_stream_seed(seed, a), p0))
        lat.upload(np.full((r, 1, L, L), p0))
        lat.step(steps)
        stats = lat.measure()
        for j in range(r):
            assert abs(rows[k * r + j, 6] - stats[j].mean[0]) <= 1e-15 * stats[j].mean[0] and rows[k * r + j, 7] == stats[j].active[0]   # setprecision(16)


def test_frame_capture_runner_and_frameDP_driver(host, tmp_path):
    """frames_expanded_percolationDornic_2D (P/:2678): every replicate's lattice at the step whose accumulated t falls in the window of
    t_event, appended to one CSV; equal (to the file's six digits) to an independent run of the same trials through the C ABI; no file when
    the window coincides with a measurement bin (the reference's dangling `else if`, P/:2763-2791).  drivers/frameDP end to end."""
    import subprocess
    L, a, dt, r, seed, t_event = 16, -0.71, 0.1, 2, 9, 1.6
    host.tmh_dp_frames.argtypes = [C.c_int] + [C.c_double] * 7 + [C.c_int, dp, dp, C.c_int, C.c_double]
    host.tmh_stream_seed.restype = C.c_ulonglong
    host.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    host.tmh_set_options(0, seed, 0, str(tmp_path).encode())
    rh0 = np.abs(np.random.default_rng(3).normal(1.0, 0.25, L * L))
    bins = np.zeros(64)
    nb = O.oracle().tmo_logarithm10_time_bins(C.c_double(20.0), C.c_double(dt), O.ptr(bins), 64)
    assert not np.any(np.abs(bins[:nb] - t_event) < dt / 2)            # 1.6 is not one of the log10 bins at dt = 0.1
    assert host.tmh_dp_frames(L, a, -2.0, 1.0, 6.25, 1.0, dt, 2.5, r, O.ptr(rh0), O.ptr(bins), nb, t_event) == 0
    path = tmp_path / "Data" / "SPDP" / "Frames" / "1RBOKCHOYFRAME_P_c_G_16_T_1.6_dt_0.1_a1_-0.71_Sig_1_dx_2.5_R_2.csv"
    assert path.exists(), os.listdir(tmp_path / "Data" / "SPDP" / "Frames")
    lines = path.read_text().splitlines()
    assert len(lines) == r * (1 + L * L) and lines[0] == " a, c, t , x,  Rho(x; t)" and lines[1 + L * L] == lines[0]
    steps, t = 0, 0.0
    while not (t_event - dt / 2.0 <= t < t_event + dt / 2.0):
        t += dt
        steps += 1
    lat = tm.Lattice(tm.dp_params(L, a=a, b=-2.0, D=6.25, sigma=1.0, dt=dt, dx=2.5, c=1.0), r)
    lat.seed(host.tmh_stream_seed(seed, a))
    lat.upload(np.tile(rh0, (r, 1)).reshape(r, 1, L, L))
    lat.step(steps)
    want = lat.download().reshape(r, L * L)
    for j in range(r):
        blk = np.array([[float(x) for x in ln.split(",")] for ln in lines[j * (1 + L * L) + 1:(j + 1) * (1 + L * L)]])
        assert np.all(blk[:, 0] == a) and np.all(blk[:, 1] == 1) and np.allclose(blk[:, 2], t_event) and np.array_equal(blk[:, 3], np.arange(L * L))
        np.testing.assert_allclose(blk[:, 4], want[j], rtol=5e-6, atol=1e-300)
    # hidden by a measurement bin: 1.5 is one of them
    assert np.any(np.abs(bins[:nb] - 1.5) < dt / 2)
    assert host.tmh_dp_frames(L, a, -2.0, 1.0, 6.25, 1.0, dt, 2.5, r, O.ptr(rh0), O.ptr(bins), nb, 1.5) == 0
    assert not (tmp_path / "Data" / "SPDP" / "Frames" / "1RBOKCHOYFRAME_P_c_G_16_T_1.5_dt_0.1_a1_-0.71_Sig_1_dx_2.5_R_2.csv").exists()
    # the driver: two a-values from one random frame
    exe = os.path.join(ROOT, "drivers", "frameDP")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    root = tmp_path / "drv"
    env = dict(os.environ, TM_DATA_ROOT=str(root), TM_SEED="7")
    out = subprocess.run([exe], input="0.1\n20\n1.6\n16\n-0.72\n-0.70\n2\n2\n", text=True, capture_output=True, env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    names = sorted(os.listdir(root / "Data" / "SPDP" / "Frames"))
    assert set(names) == {"1RBOKCHOYFRAME_P_c_G_16_T_1.6_dt_0.1_a1_-0.7_Sig_1_dx_2.5_R_2.csv", "1RBOKCHOYFRAME_P_c_G_16_T_1.6_dt_0.1_a1_-0.72_Sig_1_dx_2.5_R_2.csv"}
    assert len((root / "Data" / "SPDP" / "Frames" / names[0]).read_text().splitlines()) == r * (1 + L * L)
