"""drivers/dp_first_order_delta (BASELINE config 1's program, Percolation_FD/1storderDP_delta.cpp) against the output tree of the
UNMODIFIED reference program for the same stdin protocol (tests/golden/dp_driver_tree.json, made by tests/golden/make_dp_prelim_golden.py):
the final table AND the PRELIMINARY_* tables expanded_percolationDornic_2D dumps after the third replicate (P/:2101-2128) -- the same
files, headers, line counts, (a, c, L, t) columns and last-column rule (the replicate index while t <= 1, then the survivors so far).
The densities come from random streams and are compared by the ensemble tests."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "dp_driver_tree.json")))
EXE = os.path.join(ROOT, "drivers", "dp_first_order_delta")


@pytest.mark.gpu
def test_dp_driver_writes_the_reference_tree_with_preliminary_tables(tmp_path):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    env = dict(os.environ, TM_DATA_ROOT=str(tmp_path), TM_SEED="5")
    out = subprocess.run([EXE], input="\n".join(GOLD["stdin"]) + "\n", text=True, capture_output=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    got = {}
    for dirpath, _, fs in os.walk(str(tmp_path)):
        for f in fs:
            got[os.path.relpath(os.path.join(dirpath, f), str(tmp_path))] = open(os.path.join(dirpath, f)).read().splitlines()
    assert sorted(got) == sorted(GOLD["files"])
    for name, want in GOLD["files"].items():
        lines = got[name]
        assert len(lines) == want["lines"] and lines[0] == want["header"], name
        rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])
        assert np.isfinite(rows).all()
        assert np.array_equal(rows[:, :3], np.array(want["a_c_L"])), name
        assert np.allclose(rows[:, 3], np.array(want["t"]), rtol=0, atol=1e-12), name
        if want["last_column"] is not None:
            # while t <= 1 the last column is the replicate index j = 2 (P/:2117), then the number of the first three replicates alive in
            # that bin (P/:2123): integers in 0..3, never increasing
            ref_last = np.array(want["last_column"])
            n_early = int(np.sum(ref_last == 2.0)) if ref_last[0] == 2.0 else 0
            assert n_early == 11 and np.all(rows[:n_early, -1] == 2.0)
            tail = rows[n_early:, -1]
            assert np.all(tail == np.round(tail)) and tail.max() <= 3 and tail.min() >= 0 and np.all(np.diff(tail) <= 0)
            # three replicates are in the table: the survivor column of the running statistics says so where all are alive
            assert np.all(rows[:n_early, 6] == 3.0)
