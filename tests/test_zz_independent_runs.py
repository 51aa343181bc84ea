"""Two runs of a driver program are independent, as the reference's are (it seeds every replicate from std::random_device and puts two
random tokens into its file names, jID_<n> R/:3634 and _ID_<x>_ R/:4343): without TM_SEED a driver draws a fresh seed, prints it, and
derives the tokens from it, so a second run into the same folders leaves its frames and tables BESIDE the first run's -- the layout the
unmodified reference program produces when run twice (profiles/r02_driver_tree_comparison.txt).  With the printed seed (and
TM_FILE_IDS=1) the run repeats byte for byte; with TM_SEED alone the tokens are the constant 0 (what the other tests rely on)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "drivers", "order_1stoc_test_rietkerk")
ARGS = ["0.1", "3", "24", "2", "0.05", "0.06", "2", "10000", "TEST"]


def tree(root):
    out = {}
    for dirpath, _, fs in os.walk(os.path.join(str(root), "Data")):
        for f in fs:
            p = os.path.join(dirpath, f)
            out[os.path.relpath(p, str(root))] = open(p, "rb").read()
    return out


def run(root, **env_over):
    env = {k: v for k, v in os.environ.items() if k not in ("TM_SEED", "TM_FILE_IDS")}
    env.update(TM_DATA_ROOT=str(root), **env_over)
    out = subprocess.run([EXE] + ARGS, text=True, capture_output=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr
    return out.stdout


@pytest.mark.gpu
def test_two_runs_are_independent_and_a_printed_seed_repeats_a_run(tmp_path):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    a = tmp_path / "a"
    s1 = re.search(r"TM_SEED=(\d+)", run(a))
    t1 = tree(a)
    s2 = re.search(r"TM_SEED=(\d+)", run(a))
    t2 = tree(a)
    assert s1 and s2 and s1.group(1) != s2.group(1)
    frames1 = [k for k in t1 if "/FRAME_" in k]
    frames2 = [k for k in t2 if "/FRAME_" in k]
    assert len(frames1) == 4 and len(frames2) == 8 and set(frames1) < set(frames2)          # nothing of run 1 was overwritten or renamed
    assert all(t2[k] == t1[k] for k in frames1)
    assert len({re.search(r"_jID_(\d+)_", k).group(1) for k in frames2}) >= 3                 # one token per (run, a-value)
    assert len([k for k in t2 if "PRELIM_AGGRAND" in k]) == 2 * len([k for k in t1 if "PRELIM_AGGRAND" in k])
    # the printed seed repeats run 1 byte for byte (names included)
    b = tmp_path / "b"
    assert "TM_SEED=" not in run(b, TM_SEED=s1.group(1), TM_FILE_IDS="1")
    assert tree(b) == t1
    # TM_SEED alone: reproducible streams, constant tokens
    c = tmp_path / "c"
    run(c, TM_SEED="7")
    assert all("_jID_0_" in k for k in tree(c) if "/FRAME_" in k)
