"""bench.py's extra sections (other_configs, fke_movement_L4096, spb2_step_L4096, slab_L16384) must never cost the headline: each
runs through run_guarded() -- result, error or a watchdog that prints the one JSON line and leaves with exit code 0.  No GPU here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_result_error_and_key_selection():
    import bench
    out = {"value": 1.0}
    assert bench.run_guarded(out, 0, None, "a", 30, lambda: {"x": 1, "y": 2}) is True
    assert out["a"] == {"x": 1, "y": 2}
    assert bench.run_guarded(out, 0, None, "b", 30, lambda: {"x": 1, "y": 2}, keep=("y",)) is True
    assert out["b"] == {"y": 2}

    def boom():
        raise ValueError("no such thing")
    assert bench.run_guarded(out, 0, None, "c", 30, boom) is False
    assert out["c"] == {"error": "ValueError: no such thing"} and out["value"] == 1.0
    # a rank other than 0 keeps nothing (rank 0 prints the line)
    out1 = {}
    assert bench.run_guarded(out1, 1, None, "a", 30, lambda: None, keep=("value",)) is True and out1 == {}


def test_watchdog_prints_the_line_and_exits_zero():
    code = ("import sys, time; sys.path.insert(0, %r); import bench; out = {'metric': 'm', 'value': 2.5}; "
            "bench.run_guarded(out, 0, None, 'slow', 1, lambda: time.sleep(60)); print('NOT REACHED')" % ROOT)
    pr = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert pr.returncode == 0, pr.stderr
    lines = [l for l in pr.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 2.5 and "did not finish within 1 s" in d["slow"]["error"]


def test_default_run_names_every_baseline_config():
    """the section keys the default run carries exist in the source the driver executes"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"other_configs"', '"fke_movement_L4096"', '"spb2_step_L4096"', '"slab_L16384"', "rietkerk_L%dx%d", "rietkerk_fpe_L%dx%d", "dp_L%dx%d", "rietkerk_L%dx%d_fringe"):
        assert key in src, key
