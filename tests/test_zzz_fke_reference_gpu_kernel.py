"""The product's FKE movement step against the reference's OWN CUDA kernel on the same device buffers.

updateMovement (Rietkerk_FPE/kernel_advdiff_FKE2D.cu:92-180) is the reference's only GPU code: one thread per source site scattering
FP64 atomics into four rectangles.  oracle/Makefile compiles that file UNMODIFIED for sm_100a (-DSPB=2: one moving species) into
oracle/_ref/libtm_ref_fke_gpu.so behind an extern "C" harness (oracle/ref_harness_fke_gpu.cu: constant-memory set-up, then the
reference's per-step choreography F/:4282-4299 -- Rho_tsar zeroed, one launch).  The product's gather-form kernels are driven through
the same seam the reference's host code calls (copyToDeviceConstantMemory_AdvTerms / launchAdvDiffKernel_MultiSp of libtm_host.so).
The two sum the same non-negative terms in different orders (the reference's atomics in no fixed order at all): <= 1e-12 of the
largest value, the tolerance of the golden-vector tests of this path (tests/test_fke_kernels.py).

Runs last (zzz): it is the one -m gpu test that executes a kernel which is not the product's."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke_gpu.so")


@pytest.fixture(scope="module")
def host():
    h = C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))
    h.tmh_fke_seam_1sp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 6
    return h


@pytest.fixture(scope="module")
def ref_gpu():
    if not os.path.exists(REF_GPU):
        pytest.skip("oracle/_ref/libtm_ref_fke_gpu.so is not built (it needs /root/reference at build time)")
    r = C.CDLL(REF_GPU)
    r.ref_fke_gpu_setup.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    r.ref_fke_gpu_move.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return r


SHAPES = [
    (96, 2, 1, 3, 0.82, 1.4, 0.3),      # the seam test's shape: rows +-1, columns +-3
    (257, 1, 1, 1, 0.82, 1.2, 0.0),     # the SPB = 2 driver's windows, odd L, every site live
    (128, 3, 2, 2, 1.7, 2.0, 0.9),      # displacements up to +-3 cells, sparse field
    (64, 1, 1, 1, 40.0, 1.5, 0.5),      # displacements of +-60 cells: every copy wraps around the lattice
]


def inputs(L, bD, bvx, bvy, vmax, dead):
    rng = np.random.default_rng(L + bD)
    rho = rng.gamma(2.0, 3.0, L * L)
    rho[rng.random(L * L) < dead] = 0.0
    gamma = rng.random(L * L)
    veff = (rng.random((L * L, 2)) * 2 - 1) * vmax
    return rho, gamma, veff, O.fke_gaussian(1.0, bD), O.fke_gaussian(1.0, max(bvx, bvy))


def cuda_kernel_restated(L, bD, bvx, bvy, pref, rho, gamma, veff, gD, gV, eps=1e-10):
    """updateMovement's index arithmetic and weights for one species (kernel_advdiff_FKE2D.cu:106-176) in numpy: the skip rule :111,
    the diffusion window +-bD :115-116, mu = (int)(round(c + v prefactor) + L) % L :119-121, the advection windows +-.x on rows and +-.y
    on columns :126-127, and the four scatter loops :130-176 (np.add.at in place of the atomics)."""
    out = np.zeros(L * L)
    i = np.arange(L * L)
    live = ~(rho < eps)
    ci, cj, r, g = (i // L)[live], (i % L)[live], rho[live], gamma[live]
    mux = (np.round(ci + veff[live, 0] * pref).astype(np.int64) + L) % L
    muy = (np.round(cj + veff[live, 1] * pref).astype(np.int64) + L) % L

    def scatter(p, q, w):
        np.add.at(out, ((p + L) % L) * L + (q + L) % L, w)
    for dp in range(-bD, bD + 1):
        for dq in range(-bD, bD + 1):
            scatter(ci + dp, cj + dq, g * g * r * (gD[abs(dp)] * gD[abs(dq)]))
        for dq in range(-bvy, bvy + 1):
            scatter(ci + dp, muy + dq, g * (1 - g) * r * (gD[abs(dp)] * gV[abs(dq)]))
    for dp in range(-bvx, bvx + 1):
        for dq in range(-bD, bD + 1):
            scatter(mux + dp, cj + dq, (1 - g) * g * r * (gV[abs(dp)] * gD[abs(dq)]))
        for dq in range(-bvy, bvy + 1):
            scatter(mux + dp, muy + dq, (1 - g) * (1 - g) * r * (gV[abs(dp)] * gV[abs(dq)]))
    return out


@pytest.mark.parametrize("L,bD,bvx,bvy,pref,vmax,dead", SHAPES)
def test_emulated_product_and_oracle_match_the_cuda_kernel_restated(L, bD, bvx, bvy, pref, vmax, dead):
    """No GPU: the shapes of the GPU test below, with the CUDA kernel's arithmetic restated in numpy standing in for the kernel --
    the oracle (what the product is compared with everywhere else) and the product's kernels on host threads both equal it."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libtm_emu.so"))
    from test_fke_kernels import emu_move
    rho, gamma, veff, gD, gV = inputs(L, bD, bvx, bvy, vmax, dead)
    want = cuda_kernel_restated(L, bD, bvx, bvy, pref, rho, gamma, veff, gD, gV)
    orc = O.fke_move(L, bD, (bvx, bvx), (bvy, bvy), pref, 1e-10, rho, gamma, veff, gD, gV)
    got = emu_move(emu, L, bD, (bvx, bvx), (bvy, bvy), pref, 1e-10, rho, gamma, veff, gD, gV)
    got = got[0] if isinstance(got, tuple) else got
    assert np.abs(orc - want).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("L,bD,bvx,bvy,pref,vmax,dead", SHAPES)
def test_product_matches_reference_cuda_kernel(host, ref_gpu, L, bD, bvx, bvy, pref, vmax, dead):
    import torch
    rho, gamma, veff, gD, gV = inputs(L, bD, bvx, bvy, vmax, dead)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v).to(dev) for k, v in dict(rho=rho, gamma=gamma, veff=veff, gD=gD, gV=gV).items()}
    ours = torch.full((L * L,), 7.0, dtype=torch.float64, device=dev)      # the product's seam does not need it zeroed
    theirs = torch.empty(L * L, dtype=torch.float64, device=dev)           # the harness zeroes it, as the reference's caller does
    torch.cuda.synchronize()
    assert host.tmh_fke_seam_1sp(L, bD, bvx, bvy, pref, t["rho"].data_ptr(), ours.data_ptr(), t["gamma"].data_ptr(),
                                 t["veff"].data_ptr(), t["gD"].data_ptr(), t["gV"].data_ptr()) == 0
    assert ref_gpu.ref_fke_gpu_setup(bD, bvx, bvy, pref) == 0
    ms, msk = C.c_float(0), C.c_float(0)
    assert ref_gpu.ref_fke_gpu_move(L, t["rho"].data_ptr(), theirs.data_ptr(), t["gamma"].data_ptr(), t["veff"].data_ptr(),
                                    t["gD"].data_ptr(), t["gV"].data_ptr(), 1, C.byref(ms), C.byref(msk)) == 0
    torch.cuda.synchronize()
    got, want = ours.cpu().numpy(), theirs.cpu().numpy()
    assert np.isfinite(want).all() and want.max() > 0
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # mass: both conserve the live sites' mass up to the stencils' normalisation (the same for both)
    assert abs(got.sum() - want.sum()) <= 1e-11 * want.sum()
