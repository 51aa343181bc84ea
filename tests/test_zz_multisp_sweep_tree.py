"""Output trees of the multi-species a-sweeps with the runners' CSV side outputs (SURVEY 8 f1 + f3, e1): the SPB = 2 sweep of
Rietkerk_FD (first_order_critical_exp_delta_stochastic_MultiSp, R/:4379 with -DSPB=2) dealt over one worker and over several writes the
same bytes -- final table, FRAME_* / GAMMA_* snapshots, PRELIM_TSERIES_* / PRELIM_MOVSERIES_* checkpoints and PRELIM_AGGRAND_* tables --
and the tree has the reference's grammar (tests/test_side_outputs.py pins its content on the reference's own files)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import trophic_maelstorm_b200 as tm
from test_consumer_step2 import RUN2, setup

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


def _tree(root):
    out = {}
    for dirpath, _, files in os.walk(root):
        for f in files:
            p = os.path.join(dirpath, f)
            out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


def test_spb2_sweep_tree_with_side_outputs(tmp_path):
    L, r, div, t_max = 20, 2, 3, 2.0
    par = RUN2["mixed_par"].astype(np.float64)
    inits = np.stack([setup(L, 90 + j)[6] for j in range(r)])
    n = tm.lib().tm_device_count()
    h = C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_riet2_sweep.argtypes = [C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, dp, C.c_int]
    h.tmh_set_device_list.argtypes = [C.POINTER(C.c_int), C.c_int]
    trees = []
    for k, workers in enumerate([[0], list(range(n)) if n > 1 else [0, 0, 0]]):
        root = tmp_path / ("run%d" % k)
        h.tmh_set_options(0, 17, 0, str(root).encode())
        h.tmh_set_device_list((C.c_int * len(workers))(*workers), len(workers))
        assert h.tmh_riet2_sweep(L, par.ctypes.data_as(dp), div, t_max, 0.05, 0.06, r, inits.ctypes.data_as(dp), r) == 0
        trees.append(_tree(root))
    h.tmh_set_device_list(None, 0)
    t = trees[0]
    assert sorted(trees[1]) == sorted(t)
    assert not [k for k in t if trees[1][k] != t[k]]
    names = sorted(t)
    fdir = "Data/Rietkerk/Frames/Stochastic/2Sp/"
    pdir = "Data/Rietkerk/Prelims/Stochastic/2Sp/"
    frames = [k for k in names if k.startswith(fdir) and "/FRAME_P_c_DP_G_20_T_" in k]
    gammas = [k for k in names if k.startswith(fdir) and "/GAMMA_G_20_T_" in k]
    tser = [k for k in names if k.startswith(pdir) and "/TimeSeries/PRELIM_TSERIES_P_c_DP_G_20_T_" in k]
    mser = [k for k in names if k.startswith(pdir) and "/TimeSeries/PRELIM_MOVSERIES_P_c_DP_G_20_T_" in k]
    agg = [k for k in names if k.startswith(pdir) and "/PRELIM_AGGRAND_P_c_ID_0_DP_G_20_T_" in k]
    assert len(frames) == len(gammas) and len(frames) >= div * r            # at least the t = 0 frame of every replicate
    assert len(tser) == len(mser) and len(tser) >= div * r
    assert len(agg) == div * r
    assert all(re.search(r"_0\.05-0\.06_dP_10000/", k) for k in frames + tser + agg)   # run folder <a_st>-<a_end>_dP_<dP>
    one = t[frames[0]].decode().splitlines()
    assert one[0] == "a_c,  x,  P(x; t), G(x; t), W(x; t), O(x; t) " and len(one) == 1 + L * L and len(one[1].split(",")) == 6
    one = t[gammas[0]].decode().splitlines()
    assert one[0] == "a_c,  x,  GAM[P(x; t)], GAM[G(x; t)] " and len(one) == 1 + L * L
    one = t[mser[0]].decode().splitlines()
    assert one[0].startswith(" a , r, L, t , <GAM[P(x; t)]>_x") and len(one[1].split(",")) == 13   # 12 values + the trailing comma
    table = [k for k in names if k.startswith("Data/Rietkerk/Stochastic/2Sp/1stCC_Rietkerk_")]
    assert len(table) == 1


def test_spb2_argv_driver_binary(tmp_path):
    """drivers/order_2stoc_test_rietkerk: the eleven-argument form of Rietkerk_FD/order_2stoc_test_rietkerk.cpp:159-172 end to end --
    run folders tagged _Veq_<V*> (:303-309), the MFT initial frame of :245-283 through init_exprtk_randbiMFTframe's rule (R/:474-496),
    the final table in the reference's place with 14 columns."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "order_2stoc_test_rietkerk")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7")
    out = subprocess.run([exe, "0.1", "1.5", "16", "2", "0.05", "0.06", "2", "10000", "1", "1", "TEST"], text=True, capture_output=True,
                         env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    t = {k: v.decode() for k, v in _tree(tmp_path).items()}
    aij = 3.6 * 10.0 ** -6.08 * 20.0 ** -0.37
    m = 0.061609 * 20.0 ** -0.25 / 8760.0
    vstar = m / ((0.45 - m) * aij)
    tag = "TEST_0.05-0.06_dP_10000_Veq_%.5g" % vstar
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_16_T_1.5_dt_0.1_D1_0.0103_a1_0.05_a2_0.06_dx_0.1_Veq_%.5g_R_2.csv" % vstar   # set_Prefix, R/:133
    assert table in t, sorted(t)[:6]
    rows = np.array([[float(x) for x in ln.split(",")] for ln in t[table].splitlines()[2:]])
    assert rows.shape[1] == 14 and set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.all(rows[:, 2] == 16) and np.isfinite(rows).all()
    f0 = "Data/Rietkerk/Frames/Stochastic/2Sp/%s/FRAME_P_c_DP_G_16_T_0_dt_0.1_a_0.05_D1_0.0103_jID_0_dx_0.1_R_0.csv" % tag
    assert f0 in t, [k for k in sorted(t) if "FRAME" in k][:4]
    fr = np.array([[float(x) for x in ln.split(",")] for ln in t[f0].splitlines()[1:]])
    six = lambda x: float("%f" % x)                              # the reference's std::to_string round trip of its MFT constants
    G = six(7332.541) * (1 - np.exp(six(-15.5375) * (0.05 - six(1 / 24.0))))
    O = six((vstar + 5000) / ((vstar + 5000 * 0.2) * (0.2 / 24.0))) * 0.05
    veg = fr[:, 2]
    assert set(np.unique(veg)) <= {float("%.6g" % six(vstar)), 10000.0} and (veg == 10000.0).sum() <= 0.1 * len(veg)
    np.testing.assert_allclose(fr[:, 3], G, rtol=1e-5)           # frames carry six significant digits
    np.testing.assert_allclose(fr[:, 4], six(120.0) * 0.05, rtol=1e-5)
    np.testing.assert_allclose(fr[:, 5], O, rtol=1e-5)
    assert any(k.startswith("Data/Rietkerk/Prelims/Stochastic/2Sp/%s/TimeSeries/PRELIM_MOVSERIES_" % tag) for k in t)
    out = subprocess.run([exe, "0.1", "3"], text=True, capture_output=True)
    assert out.returncode == 1 and "Please enter the correct number of arguments" in out.stdout


def test_fpe_spb2_unity_driver_binary(tmp_path):
    """drivers/order_2stoc_unity_rietkerk: Rietkerk_FPE/order_2stoc_unity_rietkerk.cpp end to end -- dt fixed at the rounded advection
    time scale 0.12 whatever the first argument says (:58-62, 166), folders under 2Sp_mG_20 (set_Prefix(prefix, mGrazer) :197), the final
    table at Stochastic/1stOrderCC_Rietkerk_<prefix>_..., the deterministic Gaussian initial frame of F/:3954-3962."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "order_2stoc_unity_rietkerk")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7")
    L = 16
    out = subprocess.run([exe, "0.5", "1.5", str(L), "2", "0.05", "0.06", "2", "10000", "0.5", "1", "TEST"], text=True, capture_output=True,
                         env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    t = {k: v.decode() for k, v in _tree(tmp_path).items()}
    aij = 3.6 * 10.0 ** -6.08 * 20.0 ** -0.37
    m = 0.061609 * 20.0 ** -0.25 / 8760.0
    vstar = m / ((0.45 - m) * aij)
    tag = "TEST_0.05-0.06_dP_10000_Veq_%.5g" % vstar
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_16_T_1.5_dt_0.12_D1_0.0103_a1_0.05_a2_0.06_dx_0.1_Veq_%.5g_R_2.csv" % vstar
    assert table in t, sorted(t)[:6]
    rows = np.array([[float(x) for x in ln.split(",")] for ln in t[table].splitlines()[2:]])
    assert rows.shape[1] == 14 and set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.isfinite(rows).all()
    assert np.allclose(np.unique(np.round(rows[:, 3] / 0.12)) * 0.12, np.unique(rows[:, 3]))       # every bin is a multiple of dt = 0.12
    f0 = "Data/Rietkerk/Frames/Stochastic/2Sp_mG_20/%s/FRAME_P_c_DP_G_16_T_0_dt_0.12_a_0.05_D1_0.0103_jID_0_dx_0.1_R_0.csv" % tag
    assert f0 in t, [k for k in sorted(t) if "FRAME" in k][:4]
    fr = np.array([[float(x) for x in ln.split(",")] for ln in t[f0].splitlines()[1:]])
    i, j = np.divmod(np.arange(L * L), L)
    P = 5.0 * np.exp(-((i - L // 8) ** 2 + (j - L // 8) ** 2) / (2 * (L / 8.0) ** 2))             # amp = clow[Sp..]: {5, init_frac, 1, 1}
    G = 0.5 * np.exp(-((i - 7 * L // 8) ** 2 + (j - L // 8) ** 2) / (2 * (L / 16.0) ** 2))
    np.testing.assert_allclose(fr[:, 2], P, rtol=1e-5)
    np.testing.assert_allclose(fr[:, 3], G, rtol=1e-5, atol=1e-300)
    assert np.all(fr[:, 4] == 1.0) and np.all(fr[:, 5] == 1.0)


def test_spb3_argv_driver_binary(tmp_path):
    """drivers/order_3stoc_test_rietkerk: the twelve-argument form of Rietkerk_FD/order_3stoc_test_rietkerk.cpp:182-196 end to end -- dx = 1 km,
    folders under 3Sp_mG_20_mP_100 tagged _Geq_<G*>, the MFT initial frame from the driver's tables, an 18-column final table; an
    unsupported attack-rate scaling and a CFL violation are refused with the reference's exit codes."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "order_3stoc_test_rietkerk")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7")
    args = ["0.1", "1.5", "16", "2", "0.05", "0.06", "2", "10000", "1", "1", "1", "TEST"]
    out = subprocess.run([exe] + args, text=True, capture_output=True, env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    t = {k: v.decode() for k, v in _tree(tmp_path).items()}
    ajm = 3.6 * 10.0 ** -6.08 * 100.0 ** -0.37
    mm = 0.061609 * 100.0 ** -0.25 / 8760.0
    gstar = mm / ((0.85 - mm) * ajm)
    tag = "TEST_0.05-0.06_dP_10000_Geq_%.5g" % gstar
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_16_T_1.5_dt_0.1_D1_0.0298_a1_0.05_a2_0.06_dx_1_Geq_%.5g_R_2.csv" % gstar
    assert table in t, sorted(t)[:6]
    lines = t[table].splitlines()
    assert "<<Pr(x; t)>_x>_r" in lines[0]
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[2:]])
    assert rows.shape[1] == 18 and set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.isfinite(rows).all()
    f0 = "Data/Rietkerk/Frames/Stochastic/3Sp_mG_20_mP_100/%s/FRAME_P_c_DP_G_16_T_0_dt_0.1_a_0.05_D1_0.0298_jID_0_dx_1_R_0.csv" % tag
    assert f0 in t, [k for k in sorted(t) if "FRAME" in k][:4]
    fr = np.array([[float(x) for x in ln.split(",")] for ln in t[f0].splitlines()[1:]])
    six = lambda x: float("%f" % x)
    V = six(960000.0) * 0.05 + six(-40000.0)
    assert set(np.unique(fr[:, 2])) <= {V, 5 * V}                                       # speckles: max(dP, 5 V*) = 40000
    np.testing.assert_allclose(fr[:, 3], six(gstar), rtol=1e-5)
    np.testing.assert_allclose(fr[:, 4], six(4.49723728e+05) * (1 - np.exp(six(-1.73381257) * (0.05 - six(1 / 24.0)))), rtol=1e-5)
    np.testing.assert_allclose(fr[:, 5], six(-3.90344309) * 0.05 + six(5.18), rtol=1e-5)
    np.testing.assert_allclose(fr[:, 6], six(120.0) * 0.05, rtol=1e-5)
    g0 = f0.replace("FRAME_P_c_DP_G_16", "GAMMA_G_16")
    assert g0 in t and t[g0].splitlines()[0] == "a_c,  x,  GAM[P(x; t)], GAM[G(x; t)], GAM[Pr(x; t)] "
    bad = subprocess.run([exe] + args[:9] + ["7", "1", "TEST"], text=True, capture_output=True, env=env)
    assert bad.returncode == 1 and "Scaling factor for aij not supported" in bad.stderr
    cfl = subprocess.run([exe, "2.0"] + args[1:], text=True, capture_output=True, env=env)
    assert cfl.returncode == 2 and "CFL condition violated" in cfl.stdout


def test_fpe_spb3_unity_driver_binary(tmp_path):
    """drivers/order_3stoc_unity_rietkerk: Rietkerk_FPE/order_3stoc_unity_rietkerk.cpp end to end -- dt = min(0.12, 0.77), dx = 0.1, the
    Gaussian initial frame with amp = {500, 10 G*, 40, 4, 10}, 18-column table under Stochastic/1stOrderCC_Rietkerk_<prefix>_..."""
    import subprocess
    exe = os.path.join(ROOT, "drivers", "order_3stoc_unity_rietkerk")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="7")
    L = 16
    out = subprocess.run([exe, "0.5", "1.5", str(L), "2", "0.05", "0.06", "2", "10000", "1", "1", "1", "TEST"], text=True, capture_output=True,
                         env=env, timeout=1500)
    assert out.returncode == 0, out.stderr
    t = {k: v.decode() for k, v in _tree(tmp_path).items()}
    ajm = 3.6 * 10.0 ** -6.08 * 100.0 ** -0.37
    mm = 0.061609 * 100.0 ** -0.25 / 8760.0
    gstar = mm / ((0.85 - mm) * ajm)
    tag = "TEST_0.05-0.06_dP_10000_Geq_%.5g" % gstar
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_16_T_1.5_dt_0.12_D1_0.0298_a1_0.05_a2_0.06_dx_0.1_Geq_%.5g_R_2.csv" % gstar
    assert table in t, sorted(t)[:6]
    rows = np.array([[float(x) for x in ln.split(",")] for ln in t[table].splitlines()[2:]])
    assert rows.shape[1] == 18 and set(np.round(rows[:, 0], 6)) == {0.05, 0.06} and np.isfinite(rows).all()
    f0 = "Data/Rietkerk/Frames/Stochastic/3Sp_mG_20_mP_100/%s/FRAME_P_c_DP_G_16_T_0_dt_0.12_a_0.05_D1_0.0298_jID_0_dx_0.1_R_0.csv" % tag
    assert f0 in t, [k for k in sorted(t) if "FRAME" in k][:4]
    fr = np.array([[float(x) for x in ln.split(",")] for ln in t[f0].splitlines()[1:]])
    i, j = np.divmod(np.arange(L * L), L)
    bump = lambda amp, cx, cy, sd: amp * np.exp(-((i - cx) ** 2 + (j - cy) ** 2) / (2 * sd * sd))
    np.testing.assert_allclose(fr[:, 2], bump(500.0, L // 8, L // 8, L / 8.0), rtol=1e-5)
    np.testing.assert_allclose(fr[:, 3], bump(10 * gstar, 7 * L // 8, L // 8, L / 16.0), rtol=1e-5, atol=1e-300)
    np.testing.assert_allclose(fr[:, 4], bump(40.0, 7 * L // 8, 7 * L // 8, L / 16.0), rtol=1e-5, atol=1e-300)
    assert np.all(fr[:, 5] == 4.0) and np.all(fr[:, 6] == 10.0)
    bad = subprocess.run([exe, "0.5", "1.5", str(L), "2", "0.05", "0.06", "2", "10000", "1", "10", "50", "TEST"], text=True, capture_output=True, env=env)
    assert bad.returncode == 1 and "not supported" in bad.stderr


def test_fpe_spb1_driver_binary(tmp_path):
    """drivers/order_1stoc_test_rietkerk_fpe = Rietkerk_FPE/order_1stoc_test_rietkerk.cpp (BASELINE config 3's program at SPB = 1): the
    vegetation-only model under rietkerk_DorFPE_2D_MultiSp's rules -- Gaussian initial frame, checkpoints at the 85 / 90 / 95 / 100 % bins
    only -- beside the Rietkerk_FD program of the same arguments, which also writes the 60 % checkpoint."""
    import subprocess
    trees = {}
    for name in ("order_1stoc_test_rietkerk", "order_1stoc_test_rietkerk_fpe"):
        exe = os.path.join(ROOT, "drivers", name)
        assert os.path.exists(exe), "run __graft_entry__.build()"
        root = tmp_path / name
        env = dict(os.environ, TM_DATA_ROOT=str(root), TM_SEED="7")
        out = subprocess.run([exe, "0.1", "3", "16", "2", "0.05", "0.06", "2", "10000", "TEST"], text=True, capture_output=True, env=env, timeout=1500)
        assert out.returncode == 0, out.stderr
        trees[name] = {k: v.decode() for k, v in _tree(root).items()}
    fd, fpe = trees["order_1stoc_test_rietkerk"], trees["order_1stoc_test_rietkerk_fpe"]
    table = "Data/Rietkerk/Stochastic/1stOrderCC_Rietkerk_TEST_P_c_G_16_T_3_dt_0.1_D1_1.04e-05_a1_0.05_a2_0.06_dx_0.1_Geq_5_R_2.csv"
    assert table in fd and table in fpe
    cp = lambda t: sorted({re.search(r"_T_([0-9.]+)_dt", k).group(1) for k in t if "PRELIM_TSERIES" in k}, key=float)
    assert len(cp(fd)) == len(cp(fpe)) + 1 and set(cp(fpe)) < set(cp(fd))        # the 60 % mark is Rietkerk_FD's alone (F/:3032)
    f0 = [k for k in fpe if "/FRAME_P_c_DP_G_16_T_0_" in k and "_a_0.05_" in k and k.endswith("R_0.csv")]
    assert len(f0) == 1
    fr = np.array([[float(x) for x in ln.split(",")] for ln in fpe[f0[0]].splitlines()[1:]])
    L = 16
    i, j = np.divmod(np.arange(L * L), L)
    np.testing.assert_allclose(fr[:, 2], 5.0 * np.exp(-((i - L // 8) ** 2 + (j - L // 8) ** 2) / (2 * (L / 8.0) ** 2)), rtol=1e-5)
    assert np.all(fr[:, 3] == 1.0) and np.all(fr[:, 4] == 1.0)
    rows = np.array([[float(x) for x in ln.split(",")] for ln in fpe[table].splitlines()[2:]])
    assert rows.shape[1] == 10 and np.isfinite(rows).all()
