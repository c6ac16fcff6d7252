"""GPU tests of the C++ drop-in header and of the bench / smoke entry points."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_demo_against_brute_force(tmp_path):
    exe = tmp_path / "demo_api_test"
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "demo_api_test.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "ikd-tree_b200"), "-likd_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "ikd-tree_b200"), "-L/usr/local/cuda/lib64",
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS")


def test_smoke_entry():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    G.smoke()


def test_bench_contract_small():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--map-points", "150000"],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in out, key
    assert out["value"] > 0 and out["e2e"]["value"] > 0 and out["gpu_launches"] > 0
    assert out["roofline"]["bound"] == "hbm" and out["roofline"]["achieved"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] > 0 and out["e2e"]["d2h_bytes_per_step"] > 0
