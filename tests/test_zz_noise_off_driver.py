"""drivers/order_test_rietkerk: the noise-off Rietkerk a-sweep program (Rietkerk_FD/order_test_rietkerk.cpp) on the B200
path against the output tree of the UNMODIFIED reference program for the same argument lists (tests/golden/noise_off_driver_tree.json,
made by tests/golden/make_noise_off_driver_golden.py from the reference's own sources): the same folders, the same files -- the final
table and every FRAME_RAND_Basic_* snapshot of RK4_Wrapper_2D (R/:1959-1972) with its name grammar --, the same headers and line
counts, the same (a, r, L, t) columns.  The fields themselves start from a random speckle mask in the reference (random_device) and are
compared elsewhere (test_gpu_host_layer.py::test_noise_off_runner_rows_match_the_reference: injected frames, rows <= 1e-12)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "noise_off_driver_tree.json")))
EXE = os.path.join(ROOT, "drivers", "order_test_rietkerk")


def tree(root):
    files, dirs = {}, []
    for dirpath, ds, fs in os.walk(os.path.join(root, "Data")):
        dirs += [os.path.relpath(os.path.join(dirpath, d), root) for d in ds]
        for f in fs:
            p = os.path.join(dirpath, f)
            files[os.path.relpath(p, root)] = open(p).read().splitlines()
    return files, sorted(dirs)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small_kick", "big_kick", "small_kick_twice"])
def test_noise_off_driver_writes_the_reference_tree(tmp_path, case):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    g = GOLD[case]
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="11")
    for _ in range(2 if case.endswith("_twice") else 1):   # twice: no snapshot is overwritten, the replicate suffixes go on from the largest
        out = subprocess.run([EXE] + g["argv"], text=True, capture_output=True, env=env, timeout=600)
        assert out.returncode == 0, out.stderr
    files, dirs = tree(str(tmp_path))
    assert dirs == g["dirs"]
    assert sorted(files) == sorted(g["files"])
    for name, want in g["files"].items():
        got = files[name]
        assert len(got) == want["lines"], name
        assert got[0] == want["header"], name
        if "a_r_L_t" in want:
            cols = np.array([[float(x) for x in ln.split(",")[:4]] for ln in got[1:]])
            assert np.array_equal(cols, np.array(want["a_r_L_t"])), name
            rows = np.array([[float(x) for x in ln.split(",")] for ln in got[1:]])
            assert rows.shape[1] == 10 and np.isfinite(rows).all()
        else:
            body = np.array([[float(x) for x in ln.split(",")] for ln in got[1:]])       # a, i, P, W, O
            L = int(g["argv"][2])
            assert body.shape == (L * L, 5) and np.array_equal(body[:, 1], np.arange(L * L)) and np.isfinite(body).all()
    if case == "big_kick":
        # main() creates "..._dP_10000", the runner names "..._dP_1e+04" (setprecision(3), R/:1966): as in the reference nothing is
        # written and every snapshot costs one error line
        assert "Error: Could not open file" in out.stderr and "_dP_1e+04/FRAME_RAND_Basic_" in out.stderr


@pytest.mark.gpu
def test_noise_off_driver_protocols(tmp_path):
    out = subprocess.run([EXE, "0.1", "3"], text=True, capture_output=True)
    assert out.returncode == 1 and "correct number of arguments is 8" in out.stdout
    # the stdin form (Rietkerk_FPE/order_test_rietkerk.cpp is this program without the argv form): same tree as the argv form
    g = GOLD["big_kick"]
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="11")
    out = subprocess.run([EXE], input="\n".join(g["argv"]) + "\n", text=True, capture_output=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "Enter kick for high state" in out.stdout
    files, dirs = tree(str(tmp_path))
    assert dirs == g["dirs"] and sorted(files) == sorted(g["files"])
    # a snapshot is never overwritten: a second run bumps nothing here (no frames), the table is rewritten
    out = subprocess.run([EXE] + g["argv"], text=True, capture_output=True, env=env, timeout=600)
    assert out.returncode == 0 and sorted(tree(str(tmp_path))[0]) == sorted(g["files"])
