"""Host-layer pieces that need no GPU: the burn-in frame reader against save_frame's CSV."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_lattice.so"), mode=C.RTLD_GLOBAL)
    return C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))


def test_burnin_frame_round_trip(host, tmp_path):
    """save_frame (R/:2836) -> read_frame_csv (init_csvconstframe R/:909): 6 significant digits survive; header located by name."""
    host.tmh_save_frame.argtypes = [C.c_char_p, C.c_char_p, dp, C.c_int, C.c_int, C.c_double]
    host.tmh_save_frame.restype = None
    host.tmh_read_frame_csv.argtypes = [C.c_char_p, C.c_int, dp]
    L = 12
    frame = O.synthetic_rietkerk_frame(L, seed=3).reshape(3, -1).copy()
    host.tmh_save_frame(str(tmp_path).encode(), b"/FRAME_T_7_a_0.05_R_", O.ptr(frame), L, 4, 0.05)
    back = np.zeros_like(frame)
    assert host.tmh_read_frame_csv(str(tmp_path / "FRAME_T_7_a_0.05_R_4.csv").encode(), L, O.ptr(back)) == L * L
    np.testing.assert_allclose(back, frame, rtol=5e-6)
    assert host.tmh_read_frame_csv(str(tmp_path / "missing.csv").encode(), L, O.ptr(back)) == -1
    (tmp_path / "bad.csv").write_text("a_c,  x,  P(x; t), O(x; t) \n0.05,0,1,2\n")
    assert host.tmh_read_frame_csv(str(tmp_path / "bad.csv").encode(), L, O.ptr(back)) == -2
