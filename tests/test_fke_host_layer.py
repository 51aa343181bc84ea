"""The reference-facing host layer of the FKE movement step (host/tm_host_fke.h): the reference's own entry points
(set_global_FKE_params, precompute_gaussian, advdiff_FKE_MultiSp, and the -DBARRACUDA seam
copyToDeviceConstantMemory_AdvTerms / launchAdvDiffKernel_MultiSp) against the reference's golden outputs."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "fke_move.npz"))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    h = C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_set_options.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_char_p]
    h.tmh_fke_set_params_2sp.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, dp]
    h.tmh_fke_set_params_2sp.restype = None
    h.tmh_fke_override.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    h.tmh_fke_override.restype = None
    h.tmh_fke_precompute_gaussian.argtypes = [C.c_double, C.c_int, dp]
    h.tmh_fke_precompute_gaussian.restype = None
    h.tmh_fke_advdiff_2sp.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, dp, C.c_int, dp]
    h.tmh_fke_seam_1sp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 6
    return h


def set_driver_params(host, L):
    d1, v1, dt, dx, CI = GOLD["driver_inputs"]
    D = np.array([0.00025 / 24.0, d1, 0.00025 / 24.0, 0.025 / 24.0]); v = np.array([0.0, v1]); dta = np.array([dt, dt])
    out = np.zeros(7)
    host.tmh_fke_set_params_2sp(O.ptr(D), O.ptr(v), O.ptr(dta), dt, dx, L, int(CI), O.ptr(out))
    return out


def test_set_global_FKE_params_matches_reference(host):
    out = set_driver_params(host, 48)
    want = GOLD["driver_derived"]
    np.testing.assert_allclose(out[:4], want[:4], rtol=1e-15)
    assert list(out[4:7]) == list(want[4:7])


@pytest.mark.parametrize("name,sig,R", [("g_s1_R1", 1.0, 1), ("g_s1_R3", 1.0, 3), ("g_s07_R5", 0.7, 5), ("g_s25_R9", 2.5, 9)])
def test_precompute_gaussian_matches_reference(host, name, sig, R):
    got = np.zeros(R + 1)
    host.tmh_fke_precompute_gaussian(sig, R, O.ptr(got))
    np.testing.assert_allclose(got, GOLD[name], rtol=4e-16)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["driver", "wide", "wrap", "sparse"])
def test_advdiff_FKE_MultiSp_matches_reference(host, name):
    L, bD, f1, f2, pref, sD, sV, eps = GOLD[name + "_cfg"]
    L = int(L)
    host.tmh_set_options(0, 1, 0, None)
    set_driver_params(host, L)
    host.tmh_fke_override(int(bD), int(f1), int(f2), float(pref))
    rho, gamma, veff, gD, gV = (np.ascontiguousarray(GOLD[name + "_" + k]) for k in ("rho", "gamma", "veff", "gD", "gV"))
    got = np.zeros(L * L)
    assert host.tmh_fke_advdiff_2sp(L, O.ptr(rho), O.ptr(gamma), O.ptr(veff), O.ptr(gD), gD.size, O.ptr(gV), gV.size, O.ptr(got)) == 0
    want = GOLD[name + "_out"]
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.gpu
def test_barracuda_seam_with_device_pointers(host):
    """launchAdvDiffKernel_MultiSp with caller-owned device buffers (here torch tensors), the reference's choreography
    F/:4282-4299: upload, memset Rho_tsar, launch, read Rho_tsar back.  Window convention of the CUDA kernel (+-x, +-y)."""
    import torch
    L, bD, bvx, bvy, pref = 96, 2, 1, 3, 0.82
    rng = np.random.default_rng(3)
    rho = rng.gamma(2.0, 3.0, L * L); rho[rng.random(L * L) < 0.3] = 0.0
    gamma = rng.random(L * L); veff = (rng.random((L * L, 2)) * 2 - 1) * 1.4
    gD, gV = O.fke_gaussian(1.0, bD), O.fke_gaussian(1.0, max(bvx, bvy))
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v).to(dev) for k, v in dict(rho=rho, gamma=gamma, veff=veff, gD=gD, gV=gV).items()}
    tsar = torch.full((L * L,), 7.0, dtype=torch.float64, device=dev)   # need not be zeroed
    torch.cuda.synchronize()
    rc = host.tmh_fke_seam_1sp(L, bD, bvx, bvy, pref, t["rho"].data_ptr(), tsar.data_ptr(), t["gamma"].data_ptr(),
                               t["veff"].data_ptr(), t["gD"].data_ptr(), t["gV"].data_ptr())
    assert rc == 0
    got = tsar.cpu().numpy()
    want = O.fke_move(L, bD, (bvx, bvx), (bvy, bvy), pref, 1e-10, rho, gamma, veff, gD, gV)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
