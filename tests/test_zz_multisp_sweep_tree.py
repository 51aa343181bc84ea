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
