"""FKE movement kernels (csrc/tm_fke.cuh).

CPU: the product's kernel SOURCE on host threads (tests/emu) against the oracle and the reference's golden outputs.
GPU (-m gpu): the same through the C ABI (include/tm_fke.h), plus full-size properties (mass conservation,
linearity, translation covariance) at L = 4096."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "fke_move.npz"))
dp = C.POINTER(C.c_double)
TOL = 1e-12   # north_star: FP64 fields agree to <= 1e-12 relative


def case(name):
    L, bD, f1, f2, pref, sD, sV, eps = GOLD[name + "_cfg"]
    return dict(L=int(L), bD=int(bD), f1=int(f1), f2=int(f2), pref=float(pref), eps=float(eps),
                rho=GOLD[name + "_rho"], gamma=GOLD[name + "_gamma"], veff=GOLD[name + "_veff"],
                gD=GOLD[name + "_gD"], gV=GOLD[name + "_gV"], out=GOLD[name + "_out"])


def rel_err(got, want):
    return np.abs(got - want).max() / np.abs(want).max()


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libtm_emu.so"))


def emu_move(emu, L, bD, rows, cols, pref, eps, rho, gamma, veff, gD, gV):
    terms = (C.c_int * 6)(L, bD, rows[0], rows[1], cols[0], cols[1])
    out = np.empty(L * L); md = (C.c_int * 2)()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (rho, gamma, veff, gD, gV)]
    rc = emu.emu_fke_move(terms, C.c_double(pref), C.c_double(eps), *[a.ctypes.data_as(dp) for a in arrs], out.ctypes.data_as(dp), md)
    assert rc == 0
    return out, (md[0], md[1])


def random_inputs(L, seed, vmax, dead=0.3):
    rng = np.random.default_rng(seed)
    rho = rng.gamma(2.0, 3.0, L * L); rho[rng.random(L * L) < dead] = 0.0
    gamma = rng.random(L * L)
    veff = (rng.random((L * L, 2)) * 2 - 1) * vmax
    return rho, gamma, veff


@pytest.mark.parametrize("name", ["driver", "wide", "wrap", "sparse"])
def test_emulated_kernels_match_reference_golden(emu, name):
    c = case(name)
    got, md = emu_move(emu, c["L"], c["bD"], (c["f1"], c["f2"]), (c["f1"], c["f2"]), c["pref"], c["eps"],
                       c["rho"], c["gamma"], c["veff"], c["gD"], c["gV"])
    assert rel_err(got, c["out"]) <= TOL
    assert np.array_equal(got == 0.0, c["out"] == 0.0)
    assert max(md) <= int(np.ceil(np.abs(c["veff"]).max() * c["pref"])) + 1


@pytest.mark.parametrize("L,bD,rows,cols,pref,vmax", [
    (150, 2, (1, 1), (3, 3), 0.82, 0.9),     # the CUDA path's convention: rows +-x, columns +-y (.cu:126-127), two column tiles
    (33, 5, (0, 0), (0, 0), 0.0, 0.0),       # no advection: displacement 0, V window of one cell
    (16, 0, (2, 1), (1, 2), 6.5, 1.0),       # displacements up to half the lattice
    (5, 3, (3, 3), (3, 3), 1.0, 2.0),        # windows wider than the lattice: targets hit more than once per term
])
def test_emulated_kernels_match_oracle(emu, L, bD, rows, cols, pref, vmax):
    rho, gamma, veff = random_inputs(L, 7 * L + bD, vmax)
    gD = O.fke_gaussian(1.3, bD); gV = O.fke_gaussian(0.9, max(rows + cols))
    want = O.fke_move(L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)
    got, _ = emu_move(emu, L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)
    assert rel_err(got, want) <= TOL


def test_emulated_kernels_refuse_displacements_beyond_the_bound(emu):
    """kFkeMaxDisp (tm_fke.cuh): one site displaced by 70 cells makes the column pass refuse -- the moved field is NaN everywhere (loud),
    nothing is gathered (the pass would visit (2 R + 1)^2 candidate sources per target); 64 cells are still served and equal the oracle."""
    L, bD, rows, cols, pref = 200, 1, (1, 1), (1, 1), 1.0
    rho, gamma, veff = random_inputs(L, 11, 0.4)
    rho[5 * L + 7] = 3.0
    gD = O.fke_gaussian(1.0, bD); gV = O.fke_gaussian(1.0, 1)
    veff[5 * L + 7] = (70.0, 0.0)
    got, md = emu_move(emu, L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)
    assert md[0] == 70 and np.isnan(got).all()
    veff[5 * L + 7] = (64.0, -64.0)
    got, md = emu_move(emu, L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)
    assert md == (64, 64)
    assert rel_err(got, O.fke_move(L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)) <= TOL


def test_abi_exports_fke_symbols():
    lib = tm.lib()
    for name in tm.FKE_ABI_SYMBOLS:
        getattr(lib, name)
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "tm_fke.h")).read()
    import re
    declared = set(re.findall(r"^(?:int|void \*|const char \*)\s*(tm_fke_[a-z_]+)\(", hdr, re.M))
    assert declared == set(tm.FKE_ABI_SYMBOLS)


# ------------------------------------------------------------------------------------------------ GPU
def gpu_mover(L, bD, rows, cols, pref, eps, gD, gV, n_species=1):
    m = tm.FkeMover(L, n_species)
    for s in range(n_species):
        m.set_terms(s, bD, rows, cols, pref, eps, gD, gV)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["driver", "wide", "wrap", "sparse"])
def test_gpu_matches_reference_golden(name):
    c = case(name)
    m = gpu_mover(c["L"], c["bD"], (c["f1"], c["f2"]), (c["f1"], c["f2"]), c["pref"], c["eps"], c["gD"], c["gV"])
    got = m.move(c["rho"], c["gamma"], c["veff"])
    assert rel_err(got, c["out"]) <= TOL
    assert np.array_equal(got == 0.0, c["out"] == 0.0)
    again = m.move(c["rho"], c["gamma"], c["veff"])
    assert np.array_equal(got, again), "fixed summation order: bit-reproducible run to run"


@pytest.mark.gpu
@pytest.mark.parametrize("L,bD,rows,cols,pref,vmax", [
    (150, 2, (1, 1), (3, 3), 0.82, 0.9), (33, 5, (0, 0), (0, 0), 0.0, 0.0), (16, 0, (2, 1), (1, 2), 6.5, 1.0),
    (5, 3, (3, 3), (3, 3), 1.0, 2.0), (512, 9, (6, 6), (6, 6), 2.0, 1.5),
])
def test_gpu_matches_oracle(L, bD, rows, cols, pref, vmax):
    rho, gamma, veff = random_inputs(L, 7 * L + bD, vmax)
    gD = O.fke_gaussian(1.3 if bD < 9 else 3.0, bD); gV = O.fke_gaussian(0.9 if max(rows) < 6 else 2.0, max(rows + cols))
    want = O.fke_move(L, bD, rows, cols, pref, 1e-10, rho, gamma, veff, gD, gV)
    got = gpu_mover(L, bD, rows, cols, pref, 1e-10, gD, gV).move(rho, gamma, veff)
    assert rel_err(got, want) <= TOL


@pytest.mark.gpu
def test_gpu_two_species_and_errors():
    L = 64
    rho, gamma, veff = random_inputs(L, 5, 1.0)
    rho2, gamma2, veff2 = random_inputs(L, 6, 2.0)
    gD, gV = O.fke_gaussian(1.0, 1), O.fke_gaussian(1.0, 1)
    gD2, gV2 = O.fke_gaussian(2.0, 4), O.fke_gaussian(1.5, 3)
    m = tm.FkeMover(L, 2)
    with pytest.raises(tm.TmError):
        m.move(np.stack([rho, rho2]), np.stack([gamma, gamma2]), np.stack([veff, veff2]))   # terms not set
    m.set_terms(0, 1, (1, 1), (1, 1), 0.82, 1e-10, gD, gV)
    m.set_terms(1, 4, (3, 3), (3, 3), 1.7, 1e-10, gD2, gV2)
    with pytest.raises(tm.TmError):
        m.set_terms(1, 4, (5, 3), (3, 3), 1.7, 1e-10, gD2, gV2)                             # stencil too short for the window
    got = m.move(np.stack([rho, rho2]), np.stack([gamma, gamma2]), np.stack([veff, veff2]))
    want0 = O.fke_move(L, 1, (1, 1), (1, 1), 0.82, 1e-10, rho, gamma, veff, gD, gV)
    want1 = O.fke_move(L, 4, (3, 3), (3, 3), 1.7, 1e-10, rho2, gamma2, veff2, gD2, gV2)
    assert rel_err(got[0], want0) <= TOL and rel_err(got[1], want1) <= TOL
    assert m.last_displacement(1)[0] >= m.last_displacement(0)[0]


@pytest.mark.gpu
def test_gpu_fullsize_properties():
    """L = 4096 (BASELINE's largest single-GPU lattice): mass conservation, linearity in rho, translation covariance."""
    L = 4096
    rng = np.random.default_rng(99)
    rho = rng.gamma(2.0, 3.0, L * L); rho[rng.random(L * L) < 0.4] = 0.0
    gamma = rng.random(L * L)
    veff = (rng.random((L * L, 2)) * 2 - 1) * 1.2
    gD, gV = O.fke_gaussian(1.0, 3), O.fke_gaussian(1.0, 2)
    m = gpu_mover(L, 3, (2, 2), (2, 2), 0.82, 1e-10, gD, gV)
    out = m.move(rho, gamma, veff)
    assert np.all(out >= 0)
    live = rho[rho >= 1e-10].sum()
    assert abs(out.sum() - live) <= 1e-11 * live                      # stencils sum to one over symmetric windows
    out2 = m.move(2.0 * rho, gamma, veff)
    assert np.array_equal(out2, 2.0 * out)                            # linear in rho (scaling by 2 is exact)
    sh = (37, 1001)
    roll = lambda a, tail=(): np.roll(a.reshape((L, L) + tail), sh, axis=(0, 1)).reshape((L * L,) + tail)
    out3 = m.move(roll(rho), roll(gamma), roll(veff, (2,)))
    assert np.array_equal(out3, roll(out))                            # periodic lattice: no preferred origin
    ms, n = m.last_ms()
    assert n == 3 and ms > 0
