"""The C++ host layer (trophic_maelstorm_b200/host: the reference's runner signatures over the C ABI)
on the GPU: (1) its bookkeeping reproduces the reference's row assembly exactly, (2) its ensembles
agree with ensembles of the unmodified reference (tests/golden/ens_*.npz) within confidence bounds."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    tm.lib()  # libtm_lattice.so first (libtm_host.so links against it)
    h = C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_dp_run.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_int, dp, dp, C.c_int, dp, C.c_int]
    h.tmh_riet_run.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.c_int, dp, C.c_int, C.c_char_p, dp, C.c_int]
    h.tmh_dp_sweep.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_double] * 9 + [C.c_int]
    h.tmh_stream_seed.restype = C.c_ulonglong
    h.tmh_stream_seed.argtypes = [C.c_ulonglong, C.c_double]
    return h


def dp_rows(host, cubic, L, params, rho0, t_meas, r, seed):
    host.tmh_set_options(0, seed, 0, None)
    a, b, c, D, sigma, dt, dx = [float(x) for x in params]
    rows = np.zeros((4096, 9))
    tmv = np.ascontiguousarray(t_meas, dtype=np.float64)
    r0 = np.ascontiguousarray(rho0, dtype=np.float64)
    n = host.tmh_dp_run(int(cubic), L, a, b, c, D, sigma, dt, dx, r, O.ptr(r0), O.ptr(tmv), len(tmv), O.ptr(rows), 4096)
    assert n > 0, n
    off = 3 if cubic else 2
    return rows[:n, off:off + 6]  # t, mean, var, surv, active, last


def test_dp_runner_bookkeeping_matches_reference_semantics(host):
    """Same seed => same GPU trajectories; the rows must equal the reference's replicate-order
    accumulation (P/:1074-1098, var_mean_incremental_surv_runs P/:173) applied to them."""
    L, r, seed = 32, 5, 4242
    params = (1.0548, 1.0, 0.0, 0.25, 2 ** 0.5, 0.1, 1.0)
    t_meas = np.array([1.5, 2.0, 3.0, 5.0])
    rho0 = np.ones(L * L)
    rows = dp_rows(host, False, L, params, rho0, t_meas, r, seed)
    # independent replay of the same trials through the ctypes binding
    lat = tm.Lattice(tm.dp_params(L, a=params[0], b=params[1], D=params[3], sigma=params[4], dt=0.1, dx=1.0), r)
    lat.seed(host.tmh_stream_seed(seed, params[0]))      # the runner keys its streams by (TM_SEED, a): tm_host.h stream_seed
    lat.upload(np.ones((r, 1, L, L)))
    tot = int(1.0 / 0.1) + 1 + len(t_meas)
    series = np.zeros((r, tot, 2)); s = 0
    t, index = 0.0, 0
    while t < t_meas[-1] + 0.1:
        if t <= 1.0 or (index < len(t_meas) and t_meas[index] - 0.05 <= t < t_meas[index] + 0.05):
            st = lat.measure()
            for j in range(r):
                series[j, s] = (st[j].active[0], st[j].mean[0])
            s += 1
            if t > 1.0:
                index += 1
        lat.step(1)
        t += 0.1
    lat.close()
    acc = np.zeros((tot, 5))
    acc[:, 1] = series[0, :, 1]; acc[:, 4] = series[0, :, 0]; acc[:, 3] = (series[0, :, 1] > 0)
    for j in range(1, r):
        rep = np.ascontiguousarray(series[j])
        O.oracle().tmo_var_mean_incremental_surv_runs(O.ptr(acc), O.ptr(rep), tot, j)
    assert np.array_equal(rows[:, 1], acc[:, 1]) and np.array_equal(rows[:, 2], acc[:, 2])
    assert np.array_equal(rows[:, 3], acc[:, 3]) and np.array_equal(rows[:, 4], acc[:, 4])
    assert np.allclose(rows[11:, 0], t_meas)


def z_scores(ref_series, col, gpu_mean, gpu_n):
    """per-row z of (gpu ensemble mean - reference ensemble mean), surviving replicates only"""
    z = []
    for k in range(ref_series.shape[1]):
        alive = ref_series[:, k, -1] > 0
        x = ref_series[alive, k, col]
        if x.size < 8 or gpu_n[k] < 8:
            continue
        se = x.std(ddof=1) * np.sqrt(1.0 / x.size + 1.0 / gpu_n[k])
        if se == 0.0:  # deterministic row (t = 0): must agree outright
            assert abs(gpu_mean[k] - x.mean()) <= 1e-12 * abs(x.mean())
            continue
        z.append((gpu_mean[k] - x.mean()) / se)
    return np.array(z)


@pytest.mark.parametrize("name,cubic", [("ens_dp_plain", False), ("ens_dp_cubic", True)])
def test_dp_ensemble_matches_reference(host, name, cubic):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    L, r = int(g["L"]), 256
    rho0 = np.full(L * L, 0.25 if cubic else 1.0)
    rows = dp_rows(host, cubic, L, g["params"], rho0, g["t_meas"], r, seed=777)
    ref = g["series"]                                  # [n_ref][rows][t, mean, active]
    assert rows.shape[0] == ref.shape[1]
    z = z_scores(ref, 1, rows[:, 1], rows[:, 3])
    # rows are strongly correlated in time, so bound the extremes rather than count 5 % exceedances:
    # P(max |z| > 4.5) is ~1e-4 even for 60 independent rows
    assert np.abs(z).max() < 4.5 and np.abs(z).mean() < 1.6, (np.abs(z).max(), np.abs(z).mean())
    # survival probability at the last row: two-proportion z-test
    p_ref = (ref[:, -1, 2] > 0).mean(); p_gpu = rows[-1, 3] / r
    pool = (p_ref * ref.shape[0] + p_gpu * r) / (ref.shape[0] + r)
    se = max(np.sqrt(pool * (1 - pool) * (1.0 / ref.shape[0] + 1.0 / r)), 1e-9)
    assert abs(p_ref - p_gpu) / se < 4.0 or pool in (0.0, 1.0)
    if not cubic:
        # DP decay exponent delta: slope of log<rho> vs log t over t in [5, 100]; CI by bootstrap of the reference
        t = rows[:, 0]; w = (t >= 5.0) & (rows[:, 1] > 0)
        slope = lambda y: np.polyfit(np.log(t[w]), np.log(y[w]), 1)[0]
        d_gpu = slope(rows[:, 1])
        rng = np.random.default_rng(0)
        boots = []
        for _ in range(200):
            pick = rng.integers(0, ref.shape[0], ref.shape[0])
            m = np.array([ref[pick][ref[pick][:, k, 2] > 0][:, k, 1].mean() for k in range(ref.shape[1])])
            boots.append(slope(m))
        d_ref, se = np.mean(boots), np.std(boots) * np.sqrt(1 + ref.shape[0] / r)
        assert abs(d_gpu - d_ref) < 3.5 * se + 0.02, (d_gpu, d_ref, se)
        assert -0.75 < d_gpu < -0.2      # 2-D DP: delta ~ 0.45 near criticality


def test_rietkerk_ensemble_matches_reference(host):
    g = np.load(os.path.join(GOLD, "ens_rietkerk.npz"))
    L, r = int(g["L"]), 96
    ref = g["series"]                                  # [n_ref][rows][t, <W>, <O>, <P>, active]
    init = np.ascontiguousarray(np.tile(g["init"].reshape(1, -1), (r, 1)))
    tmv = np.ascontiguousarray(g["t_meas"]); ft = np.zeros(1)
    rows = np.zeros((256, 10))
    host.tmh_set_options(0, 99, 0, None)
    n = host.tmh_riet_run(L, float(g["a"]), float(g["dt"]), r, O.ptr(init), O.ptr(tmv), len(tmv), O.ptr(ft), 0, None, O.ptr(rows), 256)
    assert n == ref.shape[1]
    rows = rows[:n]                                    # a, r, L, t, <W>, <O>, <P>, var, surv, active
    nsurv = rows[:, 8]
    zP = z_scores(ref[:, :, [0, 3, 4]], 1, rows[:, 6], nsurv)
    assert np.abs(zP[1:]).max() < 4.5, np.abs(zP).max()
    # deterministic-dominated water fields: means agree to a relative 1e-4 (noise enters only through P)
    np.testing.assert_allclose(rows[:, 4], ref[:, :, 1].mean(axis=0), rtol=2e-4)
    np.testing.assert_allclose(rows[:, 5], ref[:, :, 2].mean(axis=0), rtol=2e-4)
    # active-site counts: ensemble means within bounds
    zA = z_scores(ref[:, :, [0, 4, 4]], 1, rows[:, 9], nsurv)
    assert np.abs(zA[1:]).max() < 4.5, np.abs(zA).max()


def test_dp_sweep_writes_reference_csv(host, tmp_path):
    """first_order_critical_exp_delta (P/:2422): file name grammar and column layout of the sweep CSV."""
    host.tmh_set_options(0, 5, 0, str(tmp_path).encode())
    rc = host.tmh_dp_sweep(1, 32, 2, 3.0, -0.72, -0.70, -2.0, 1.0, 6.25, 1.0, 0.1, 2.5, 2)
    assert rc == 0
    path = tmp_path / "Data" / "SPDP" / "1stOrder_P_c_Delta_DP_G_32_T_3_dt_0.1_c_1_a1_-0.72_a2_-0.7_dx_2.5_Sig_1_R_2.csv"
    assert path.exists(), list((tmp_path / "Data" / "SPDP").iterdir())
    lines = path.read_text().splitlines()
    assert lines[0] == " a , c , L,  t ,  <<Rho(x; t)>_x>_r, Var[<Rho(t)>_x]_r, # Surviving Runs, # Active Sites, # Surviving Runs Alt "
    bins = O.oracle().tmo_logarithm10_time_bins(3.0, 0.1, O.ptr(np.zeros(64)), 64)
    assert len(lines) - 1 == 2 * (11 + bins)
    first = lines[1].split(",")
    assert len(first) == 9 and float(first[0]) == -0.72 and float(first[2]) == 32 and float(first[4]) == 0.25


def test_config1_driver_binary(tmp_path):
    """drivers/dp_first_order_delta: BASELINE config 1 (L=128, a -0.72 -> -0.70, b=-2, c=1) end to end,
    same stdin protocol and output file as Percolation_FD/1storderDP_delta.cpp."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "dp_first_order_delta")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7")
    out = subprocess.run([exe], input="0.1\n20\n128\n2\n-0.72\n-0.70\n3\n", text=True, capture_output=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    files = list((tmp_path / "Data" / "SPDP").iterdir())
    assert len(files) == 1 and files[0].name.startswith("1stOrder_P_c_Delta_DP_G_128_T_20_dt_0.1_c_1_a1_-0.72_a2_-0.7_dx_2.5_Sig_1_R_2")
    rows = np.loadtxt(files[0], delimiter=",", skiprows=1)
    assert rows.shape[1] == 9 and set(np.round(rows[:, 0], 6)) == {-0.72, -0.71, -0.70}
    assert np.all(rows[:, 2] == 128) and np.all(rows[rows[:, 3] == 0][:, 4] == 0.25)
    assert np.all(rows[:, 4] > 0) and np.all(rows[:, 7] <= 128 * 128)


def test_rietkerk_sweep_writes_reference_tree(host, tmp_path):
    """first_order_critical_exp_delta_stochastic_MultiSp (R/:4379): final table, per-replicate prelim tables and
    frame snapshots in the reference's directory/file grammar and formats."""
    host.tmh_riet_sweep.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    host.tmh_set_options(0, 31, 0, str(tmp_path).encode())
    L, div, r = 32, 2, 2
    assert host.tmh_riet_sweep(L, div, 3.0, 0.05, 0.06, 0.1, r) == 0
    root = tmp_path / "Data" / "Rietkerk"
    table = root / "Stochastic" / "1Sp" / "1stCC_Rietkerk_DiC-NREF-LI_P_c_G_32_T_3_dt_0.1_D1_1.04e-05_a1_0.05_a2_0.06_dx_0.1_R_2.csv"
    assert table.exists(), list((root / "Stochastic" / "1Sp").iterdir())
    lines = table.read_text().splitlines()
    assert lines[0].startswith(" a , r, L, t , <<W(x; t)>_x>_r, <<O(x; t)>_x>_r,  <<P(x; t)>_x>_r, Var[<P(x; t)>_x]_r")
    nbins = O.oracle().tmo_logarithm10_time_bins(3.0, 0.1, O.ptr(np.zeros(64)), 64) + 1     # {0} U log bins
    rows = np.array([[float(x) for x in l.split(",")] for l in lines[2:]])
    assert rows.shape == (div * nbins, 10)
    assert set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.all(rows[:, 1] == r) and np.all(rows[:, 2] == L)
    first = rows[0]                                   # t = 0: the MFT frame, W* a = 5 a, O = 120.47 a + 0.58
    assert first[3] == 0.0 and abs(first[4] - 0.25) < 1e-12 and abs(first[5] - (120.4744181 * 0.05 + 0.58243061)) < 1e-9
    assert first[8] == r and first[9] == L * L
    # frames: t = 0 and the last ten frame_tmeas windows, 6-significant-digit CSV with the frame header
    fdir = root / "Frames" / "Stochastic" / "1Sp" / "DiC-NREF-LI_0.05-0.06_dP_10000"
    frames = sorted(p.name for p in fdir.iterdir())
    assert "FRAME_P_c_DP_G_32_T_0_dt_0.1_a_0.05_D1_1.04e-05_jID_0_dx_0.1_R_0.csv" in frames
    txt = (fdir / frames[0]).read_text().splitlines()
    assert txt[0] == "a_c,  x,  P(x; t), W(x; t), O(x; t) " and len(txt) == 1 + L * L and len(txt[1].split(",")) == 5
    # prelim tables: one per (a, replicate), trailing comma, r+1 in the second column
    pdir = root / "Prelims" / "Stochastic" / "1Sp" / "DiC-NREF-LI_0.05-0.06_dP_10000"
    prelims = sorted(p.name for p in pdir.iterdir() if p.is_file())
    assert len(prelims) == div * r and prelims[0].startswith("PRELIM_AGGRAND_P_c_ID_0_DP_G_32_T_")
    pl = (pdir / prelims[0]).read_text().splitlines()
    assert pl[1].endswith(",") and pl[1].split(",")[1] in ("1", "2") and len(pl) == 2 + nbins - 1 + 0 or len(pl) >= nbins


    # time-series checkpoints of every replicate at 60/85/90/95/100 % of the measurement bins (R/:3710, save_prelimfileswrapper R/:3029)
    ts = sorted(p.name for p in (pdir / "TimeSeries").iterdir())
    marks = sorted({int(nbins * f) for f in (0.6, 0.85, 0.9, 0.95)} | {nbins - 1})
    assert len(ts) == div * r * len(marks) and all(n.startswith("PRELIM_TSERIES_P_c_DP_G_32_T_") for n in ts)
    mine = [n for n in ts if "_dt_0.1_a_0.05_" in n and n.endswith("_R_0.csv")]
    assert len(mine) == len(marks)
    tl = max(((pdir / "TimeSeries" / n).read_text().splitlines() for n in mine), key=len)   # the checkpoint at the last bin
    trow = np.array([[float(x) for x in l.rstrip(",").split(",")] for l in tl[1:]])
    assert trow.shape == (nbins, 10) and np.all(trow[:, 8] == 1) and np.all(trow[:, 7] == 0)   # one replicate: one survivor, variance 0
    assert abs(trow[0, 4] - 0.25) < 1e-12


@pytest.mark.gpu
def test_rietkerk_runner_from_burnin_frames(host, tmp_path):
    """use_burnin_frames (init_burnin_wrapper R/:836): every replicate starts from a saved frame; the t = 0 row of the
    runner reports that frame's means."""
    host.tmh_save_frame.argtypes = [C.c_char_p, C.c_char_p, dp, C.c_int, C.c_int, C.c_double]
    host.tmh_save_frame.restype = None
    host.tmh_use_burnin.argtypes = [C.c_char_p, C.c_char_p]
    host.tmh_use_burnin.restype = None
    L = 32
    frame = O.synthetic_rietkerk_frame(L, seed=9, fringe=False).reshape(3, -1).copy()
    host.tmh_save_frame(str(tmp_path).encode(), b"/FRAME_T_76000_a_0.05_R_", O.ptr(frame), L, 0, 0.05)
    host.tmh_use_burnin(str(tmp_path).encode(), b"/FRAME_T_76000_a_0.05_R_")
    host.tmh_set_options(0, 5, 0, str(tmp_path).encode())
    tmv = np.array([0.0, 0.5, 1.0]); rows = np.zeros((64, 10)); ft = np.zeros(1)
    n = host.tmh_riet_run(L, 0.05, 0.1, 2, None, O.ptr(tmv), len(tmv), O.ptr(ft), 0, None, O.ptr(rows), 64)
    host.tmh_use_burnin(None, None)
    assert n == 3
    back = np.zeros_like(frame)
    host.tmh_read_frame_csv.argtypes = [C.c_char_p, C.c_int, dp]
    host.tmh_read_frame_csv(str(tmp_path / "FRAME_T_76000_a_0.05_R_0.csv").encode(), L, O.ptr(back))
    assert abs(rows[0, 6] - back[0].mean()) <= 1e-12 * back[0].mean() and abs(rows[0, 4] - back[1].mean()) < 1e-12


# ---- spreading experiments: tm_measure_spread + spreading_exp_Dornic_2D ---------------

def test_measure_spread_matches_oracle():
    """N(t), sum of squared distances about the seed and the border test (P/:156, P/:380): exact (integers)."""
    L, nt = 64, 5
    rng = np.random.default_rng(3)
    f = np.zeros((nt, 1, L, L))
    f[0, 0, L // 2, L // 2] = 1.0                                   # the solitary seed: N=1, R^2=0, no border
    f[1, 0] = (rng.random((L, L)) < 0.05) * rng.random((L, L))      # scattered, touches the border
    f[2, 0, 10:50, 12:40] = (rng.random((40, 28)) < 0.5) * 0.3      # interior blob
    f[3, 0, L - 1, 7] = 1e-300                                      # a single active site on the last row
    f[4, 0, 20, 0] = 2.0; f[4, 0, 5, 5] = -1.0                      # first column; a negative site is not active
    lat = tm.Lattice(tm.dp_params(L, a=1.0548, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0), nt)
    lat.upload(f)
    for (cx, cy) in ((L // 2, L // 2), (3, 60)):
        got = lat.measure_spread(cx, cy)
        for t in range(nt):
            sq, n, border = O.spread_stats(f[t, 0], L, cx, cy)
            assert (got[t].n_active, got[t].sq_dist, got[t].border) == (n, sq, border), t
    lat.close()


def test_spreading_runner_ensemble_matches_reference(host):
    """spreading_exp_Dornic_2D through the host layer (replicates as one batch of trials) against 320 replicates of the
    unmodified reference: <N(t)> and <R^2(t)> over surviving runs and P_surv(t) within 4.5 standard errors."""
    g = np.load(os.path.join(GOLD, "ens_dp_spread.npz"))
    L, tmv, ref = int(g["L"]), np.ascontiguousarray(g["t_meas"]), g["series"]     # [rep][row][t, N, R2, surv]
    a, b, D, sigma, dt, dx = [float(x) for x in g["params"]]
    r = 1024
    host.tmh_dp_spread_run.argtypes = [C.c_int] + [C.c_double] * 6 + [C.c_int, dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int]
    host.tmh_set_options(0, 99, 0, None)
    rho0 = np.zeros(L * L); rho0[L * (L // 2) + L // 2] = 1.0
    rows = np.zeros((4096, 6))
    n = host.tmh_dp_spread_run(L, a, b, D, sigma, dt, dx, r, O.ptr(rho0), O.ptr(tmv), len(tmv), L // 2, L // 2, O.ptr(rows), 4096)
    assert n == ref.shape[1], n
    rows = rows[:n]
    assert np.allclose(rows[:, 2], ref[0, :, 0])
    assert rows[0, 3] == 1.0 and rows[0, 4] == 0.0 and rows[0, 5] == r            # t = 0: the seed alone, everyone alive
    assert (np.diff(rows[:, 5]) <= 0).all()                                        # survival only decreases
    nref = ref.shape[0]
    surv_ref = ref[:, :, 3]
    p_ref, p_gpu = surv_ref.mean(0), rows[:, 5] / r
    p_pool = (p_ref * nref + p_gpu * r) / (nref + r)                              # pooled: p_ref is exactly 1 at early times
    se_p = np.sqrt(np.maximum(p_pool * (1 - p_pool), 1e-4) * (1.0 / nref + 1.0 / r))
    assert (np.abs(p_gpu - p_ref) <= 4.5 * se_p + 1e-12).all(), np.abs(p_gpu - p_ref) / se_p
    for col, gcol in ((1, 3), (2, 4)):                                             # N(t), R^2(t) over surviving runs
        for k in range(1, n):
            alive = surv_ref[:, k] > 0
            if alive.sum() < 20 or rows[k, 5] < 20:
                continue
            x = ref[alive, k, col]
            se = x.std(ddof=1) * np.sqrt(1.0 / alive.sum() + 1.0 / rows[k, 5])
            assert abs(rows[k, gcol] - x.mean()) <= 4.5 * se + 1e-9, (col, k, rows[k, gcol], x.mean(), se)


# ---- the noise-off Rietkerk runner RK4_Wrapper_2D (R/:1816-2166) -------------------------------------------
HERE = os.path.dirname(os.path.abspath(__file__))
DETRUN = np.load(os.path.join(HERE, "golden", "riet_det_runner.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["alive", "one_dead"])
def test_noise_off_runner_rows_match_the_reference(name):
    """tm_b200::RK4_Wrapper_2D (the reference's argument list) from the same initial frames as the UNMODIFIED reference runner
    (tests/golden/make_riet_det_runner_golden.py): the same rows -- measured at every step while t <= 1 and in the t_meas
    windows of the accumulated t += dt; times, surviving runs and active sites exact; means <= 1e-12; the running variance
    (a difference of nearly equal means) to its conditioning.  `one_dead`: a replicate without vegetation leaves its water
    entries at zero (the reference's `break`, R/:1953) and drops out of the surviving-run statistics."""
    host = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    dp = C.POINTER(C.c_double)
    host.tmh_riet_det_run.argtypes = [dp, C.c_int, C.c_int, dp, dp, C.c_int, dp, C.c_int]
    host.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    host.tmh_set_options(0, 1, 0, None)
    par, L, frames, tm_, want = DETRUN[name + "_par"], int(DETRUN[name + "_L"]), DETRUN[name + "_frames"], DETRUN[name + "_tmeas"], DETRUN[name + "_rows"]
    frames = np.ascontiguousarray(frames)
    rows = np.zeros((len(want) + 8, 10))
    n = host.tmh_riet_det_run(par.ctypes.data_as(dp), L, len(frames), frames.ctypes.data_as(dp), tm_.ctypes.data_as(dp), len(tm_),
                              rows.ctypes.data_as(dp), len(rows))
    assert n == len(want)
    got = rows[:n]
    assert np.array_equal(got[:, [0, 1, 2, 3, 8, 9]], want[:, [0, 1, 2, 3, 8, 9]])          # a, r, L, t, #surviving, #active
    for col in (4, 5, 6):
        assert np.abs(got[:, col] - want[:, col]).max() <= 1e-12 * np.abs(want[:, col]).max(), col
    assert np.allclose(got[:, 7], want[:, 7], rtol=5e-3, atol=1e-13 * np.abs(want[:, 6]).max() ** 2)
    # the default initial frame (no frames supplied): the reference's speckle rule around the analytic steady state
    n2 = host.tmh_riet_det_run(par.ctypes.data_as(dp), L, 2, None, tm_.ctypes.data_as(dp), len(tm_), rows.ctypes.data_as(dp), len(rows))
    assert n2 == len(want) and np.all(rows[:n2, 8] == 2) and np.isfinite(rows[:n2]).all()


@pytest.mark.gpu
def test_noise_off_sweep_writes_the_reference_csv(tmp_path):
    """first_order_critical_exp_delta (R/:2170, the noise-off a-sweep): one CSV in the reference's place, header and
    16-digit rows {a, r, L, t, <W>, <O>, <P>, var, #surviving, #active}, a-values in order, (1/dt + 1 + #bins) rows each."""
    host = C.CDLL(os.path.join(os.path.dirname(HERE), "trophic_maelstorm_b200", "libtm_host.so"))
    host.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    host.tmh_riet_det_sweep.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    host.tmh_set_options(0, 9, 0, str(tmp_path).encode())
    L, div, t_max, dt, r = 32, 2, 3.0, 0.1, 2
    assert host.tmh_riet_det_sweep(L, div, t_max, 0.05, 0.06, dt, r) == 0
    d = tmp_path / "Data" / "Rietkerk" / "Non-Stochastic"
    files = list(d.glob("1stOrderCC_Rietkerk_P_c_Delta_DP_G_32_T_3_dt_0.1_D0_*_a1_0.05_a2_0.06_dx_0.1_gmax_*_alpha_*_R_2.csv"))
    assert len(files) == 1, list(d.iterdir())
    lines = files[0].read_text().splitlines()
    assert lines[0].startswith(" a , r, L, t , <<W(x; t)>_x>_r")
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])
    n_bins = len(O.time_bins(t_max, dt)) if hasattr(O, "time_bins") else None
    per_a = len(rows) // div
    assert rows.shape[1] == 10 and len(rows) == per_a * div and (n_bins is None or per_a == int(1.0 / dt) + 1 + n_bins)
    assert np.all(rows[:per_a, 0] == 0.05) and np.all(rows[per_a:, 0] == 0.06) and np.all(rows[:, 1] == r) and np.all(rows[:, 2] == L)
    # a < 1: the steady state without vegetation plus 1.5 % speckles of dP (R/:1857, 1914); diffusion spreads them over the lattice
    assert np.all(rows[:, 8] == r) and np.all(np.diff(rows[:per_a, 3]) > 0)
    assert 0 < rows[0, 9] < 0.05 * L * L and np.all(np.diff(rows[:per_a, 9]) >= 0) and rows[per_a - 1, 9] == L * L


# ---- the ensemble scheduler (VERDICT r1 missing #1; R/:4509-4542, P/:2462-2488) -------------------------------------------
def _tree(root):
    """{relative path: bytes} of every file under root"""
    out = {}
    for dirpath, _, files in os.walk(root):
        for f in files:
            p = os.path.join(dirpath, f)
            out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


def _worker_lists():
    """one worker; then every visible GPU (on a one-GPU box: three workers sharing GPU 0, which drives the same threads,
    contexts and chunking)"""
    import trophic_maelstorm_b200 as tm
    n = tm.lib().tm_device_count()
    return [[0], list(range(n)) if n > 1 else [0, 0, 0]]


@pytest.mark.gpu
def test_sweep_outputs_do_not_depend_on_the_gpu_schedule(host, tmp_path):
    """A sweep dealt over several workers / GPUs, or cut into small device contexts, writes byte-identical output trees:
    Philox keys and initial frames hang on the global trial id, all order-dependent host arithmetic happens after the join."""
    host.tmh_riet_sweep.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    host.tmh_set_device_list.argtypes = [C.POINTER(C.c_int), C.c_int]
    host.tmh_set_devices.argtypes = [C.c_int, C.c_int, C.c_int]
    trees = []
    for k, (workers, cap) in enumerate([(_worker_lists()[0], 0), (_worker_lists()[1], 0), (_worker_lists()[1], 2)]):
        root = tmp_path / ("run%d" % k)
        host.tmh_set_options(0, 31, 0, str(root).encode())
        host.tmh_set_devices(0, 1, cap)
        host.tmh_set_device_list((C.c_int * len(workers))(*workers), len(workers))
        assert host.tmh_riet_sweep(32, 3, 3.0, 0.05, 0.06, 0.1, 3) == 0          # 3 a-values x 3 replicates = 9 trials
        assert host.tmh_dp_sweep(1, 32, 4, 3.0, -0.72, -0.70, -2.0, 1.0, 6.25, 1.0, 0.1, 2.5, 2) == 0   # 4 a-values
        trees.append(_tree(root))
    host.tmh_set_device_list(None, 0)
    host.tmh_set_devices(0, 1, 0)
    assert len(trees[0]) > 20
    for t in trees[1:]:
        assert sorted(t) == sorted(trees[0])
        diff = [k for k in t if t[k] != trees[0][k]]
        assert not diff, diff[:5]


@pytest.mark.gpu
def test_rietkerk_argv_driver_binary(tmp_path):
    """drivers/order_1stoc_test_rietkerk: the nine-argument form of Rietkerk_FD/order_1stoc_test_rietkerk.cpp:99-110 end to end
    (folders of set_Prefix R/:113-135, final table of R/:4601); 1 worker and all workers give the same bytes."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "order_1stoc_test_rietkerk")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    trees = []
    for k, workers in enumerate(_worker_lists()):
        root = tmp_path / ("d%d" % k)
        env = dict(os.environ, TM_DATA_ROOT=str(root), TM_SEED="7", TM_DEVICE_LIST=",".join(map(str, workers)))
        out = subprocess.run([exe, "0.1", "3", "32", "2", "0.05", "0.06", "2", "10000", "TEST"], text=True, capture_output=True, env=env, timeout=300)
        assert out.returncode == 0, out.stderr
        trees.append(_tree(root))
    t = trees[0]
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_32_T_3_dt_0.1_D1_1.04e-05_a1_0.05_a2_0.06_dx_0.1_Geq_5_R_2.csv"
    assert table in t, sorted(t)[:10]
    lines = t[table].decode().splitlines()
    assert lines[0].startswith(" a , r, L, t ,")
    rows = np.array([[float(x) for x in l.split(",")] for l in lines[2:]])
    assert rows.shape[1] == 10 and set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.all(rows[:, 2] == 32)
    assert any(k.startswith("Data/Rietkerk/Frames/Stochastic/1Sp/TEST_0.05-0.06_dP_10000_Geq_5/FRAME_") for k in t)
    assert any(k.startswith("Data/Rietkerk/Prelims/Stochastic/1Sp/TEST_0.05-0.06_dP_10000_Geq_5/TimeSeries/") for k in t)
    assert trees[1] == t
    out = subprocess.run([exe, "0.1", "3"], text=True, capture_output=True)
    assert out.returncode == 1 and "correct number of arguments is 9" in out.stdout
