"""FKE movement: the oracle restatement against the reference's own outputs (tests/golden/fke_move.npz,
made by tests/golden/make_fke_golden.py from oracle/_ref/libtm_ref_fke.so).  CPU only."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fke_move.npz"))
CASES = ["driver", "wide", "wrap", "sparse"]


def case(name):
    L, bD, f1, f2, pref, sD, sV, eps = GOLD[name + "_cfg"]
    return dict(L=int(L), bD=int(bD), f1=int(f1), f2=int(f2), pref=float(pref), sD=float(sD), sV=float(sV), eps=float(eps),
                rho=GOLD[name + "_rho"], gamma=GOLD[name + "_gamma"], veff=GOLD[name + "_veff"],
                gD=GOLD[name + "_gD"], gV=GOLD[name + "_gV"], out=GOLD[name + "_out"])


def test_derived_constants_match_reference():
    d1, v1, dt, dx, CI = GOLD["driver_inputs"]
    want = GOLD["driver_derived"]   # beta_vel, sigma_D, sigma_vD, prefactor, bound_D, bound_vD.first, .second, epsilon
    d = O.fke_params(d1, v1, dt, dx, int(CI))
    np.testing.assert_allclose([d.beta_vel, d.sigma_D, d.sigma_vD, d.mu_vel_prefactor], want[:4], rtol=1e-15)
    assert [d.bound_D, d.bound_vD_first, d.bound_vD_second] == [int(x) for x in want[4:7]]
    assert want[7] == 1e-10


@pytest.mark.parametrize("name,sig,R", [("g_s1_R1", 1.0, 1), ("g_s1_R3", 1.0, 3), ("g_s07_R5", 0.7, 5), ("g_s25_R9", 2.5, 9)])
def test_gaussian_stencil_matches_reference(name, sig, R):
    got = O.fke_gaussian(sig, R)
    np.testing.assert_allclose(got, GOLD[name], rtol=4e-16)   # libm exp vs the reference build's: <= 1 ulp
    assert abs(got[0] + 2 * got[1:].sum() - 1.0) < 1e-15


@pytest.mark.parametrize("name", CASES)
def test_move_matches_reference(name):
    c = case(name)
    got = O.fke_move(c["L"], c["bD"], (c["f1"], c["f2"]), (c["f1"], c["f2"]), c["pref"], c["eps"],
                     c["rho"], c["gamma"], c["veff"], c["gD"], c["gV"])
    # same loop order as the reference; its build contracts a*b+c into FMAs, the oracle does not
    np.testing.assert_allclose(got, c["out"], rtol=1e-13, atol=1e-300)
    assert np.array_equal(got == 0.0, c["out"] == 0.0)


def test_move_conserves_mass_with_symmetric_windows():
    c = case("driver")
    got = O.fke_move(c["L"], c["bD"], (c["f1"], c["f2"]), (c["f1"], c["f2"]), c["pref"], c["eps"],
                     c["rho"], c["gamma"], c["veff"], c["gD"], c["gV"])
    live = c["rho"][c["rho"] >= c["eps"]].sum()
    assert abs(got.sum() - live) <= 1e-12 * live
