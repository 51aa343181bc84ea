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


@pytest.mark.parametrize("fixture,n_fields,spb", [("side_outputs_2sp.json.gz", 4, 2), ("side_outputs_3sp.json.gz", 5, 3)])
def test_multispecies_burnin_reader_on_the_references_own_frames(host, tmp_path, fixture, n_fields, spb):
    """the -DINIT=2 path for SPB = 2 / 3 (init_burnin_wrapper R/:836, init_exprtk_readCSVcolumns_frame R/:609): FRAME_* files written by the
    UNMODIFIED reference (the side-output fixtures) are read back field by field -- columns located by header name, water last -- and the
    replicate rule picks "<pattern><r'>.csv" with r' = (replicate + offset) mod (files available)."""
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, "tests", "golden", fixture), "rt") as f:
        g = json.load(f)
    L = g["L"]
    frames = {k: v for k, v in g["files"].items() if "/FRAME_" in k and "_T_1.2_" in k}
    assert len(frames) == 2
    host.tmh_read_frame_csv_fields.argtypes = [C.c_char_p, C.c_int, C.c_int, dp]
    host.tmh_burnin_frame.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, dp]
    parsed = {}
    for k, text in frames.items():
        rep = int(k[-5])
        (tmp_path / ("FRAME_T_76000_a_0.05_R_%d.csv" % rep)).write_text(text)
        rows = np.array([[float(x) for x in ln.split(",")] for ln in text.splitlines()[1:]])
        parsed[rep] = rows[:, 2:].T.copy()                      # memory order = column order in save_frame's layout
    back = np.zeros((n_fields, L * L))
    for rep in (0, 1):
        assert host.tmh_read_frame_csv_fields(str(tmp_path / ("FRAME_T_76000_a_0.05_R_%d.csv" % rep)).encode(), L, n_fields, O.ptr(back)) == L * L
        assert np.array_equal(back, parsed[rep])
    # a 1Sp reader on the same file takes P, W, O by name; a 5-field reader on a 4-field file reports the missing column
    three = np.zeros((3, L * L))
    assert host.tmh_read_frame_csv_fields(str(tmp_path / "FRAME_T_76000_a_0.05_R_0.csv").encode(), L, 3, O.ptr(three)) == L * L
    assert np.array_equal(three[0], parsed[0][0]) and np.array_equal(three[1], parsed[0][-2]) and np.array_equal(three[2], parsed[0][-1])
    if n_fields == 4:
        assert host.tmh_read_frame_csv_fields(str(tmp_path / "FRAME_T_76000_a_0.05_R_0.csv").encode(), L, 5, O.ptr(np.zeros((5, L * L)))) == -2
    got = [np.zeros((n_fields, L * L)) for _ in range(3)]
    for j in range(3):
        assert host.tmh_burnin_frame(spb, (str(tmp_path) + "/").encode(), b"FRAME_T_76000_a_0.05_R_", L, j, O.ptr(got[j])) == 0
    assert np.array_equal(got[0], got[2]) and not np.array_equal(got[0], got[1])       # two files: replicates alternate
    assert any(np.array_equal(got[0], parsed[rep]) for rep in (0, 1)) and any(np.array_equal(got[1], parsed[rep]) for rep in (0, 1))
