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


def test_cpp_concurrent_callers_and_long_running_map(tmp_path):
    """Unmodified-caller behaviour: 16 OpenMP threads issuing single-query Nearest_Search calls (combined into shared
    launches, results bit-equal to the batched call), and bounded ids / payload on a long-running map."""
    exe = tmp_path / "concurrent_api_test"
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-fopenmp", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "concurrent_api_test.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "ikd-tree_b200"), "-likd_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "ikd-tree_b200"), "-L/usr/local/cuda/lib64",
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    r = subprocess.run([str(exe), "16"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS")


def test_smoke_entry():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    G.smoke()


BENCH_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity")


def run_bench(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_bench_contract_scanloop_small():
    out = run_bench("--workload", "scanloop", "--steps", "3", "--warmup", "3", "--map-points", "150000")
    for key in BENCH_KEYS + ("roofline_add_points", "scan_p50_ms"):
        assert key in out, key
    assert out["value"] > 0 and out["e2e"]["value"] > 0 and out["gpu_launches"] > 0
    assert out["roofline"]["bound"] == "hbm" and out["roofline"]["achieved"] > 0
    assert out["roofline_add_points"]["achieved"] > 0 and out["roofline_add_points"]["v_box_mean"] > 1
    assert out["e2e"]["h2d_bytes_per_step"] > 0 and out["e2e"]["d2h_bytes_per_step"] > 0
    assert out["parity"]["ok"] is True  # the bench checked its own answers against the CPU oracle


def test_bench_contract_largebatch_small():
    """The default workload (configs[3]) at a reduced size: same code path as the driver's run, with the reference leg."""
    out = run_bench("--steps", "2", "--warmup", "3", "--map-points", "2000000", "--queries", "4000000")
    for key in BENCH_KEYS:
        assert key in out, key
    assert out["config"]["workload"].startswith("configs[3]") and out["scaling"] == "strong"
    assert out["value"] > 0 and out["e2e"]["value"] > 0 and out["gpu_launches"] > 0
    assert out["roofline"]["kernel"].startswith("knn_reg_persist_kernel") and 0 < out["roofline"]["frac"] < 2
    assert out["cpu_baseline"]["kind"] == "reference" and out["cpu_baseline"]["value"] > 0
    assert out["parity"]["ok"] is True and out["parity"]["queries"] == 1000000
    ref = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--map-points", "2000000", "--queries", "4000000")
    assert ref["impl"] == "reference" and ref["config"] == out["config"] and ref["metric"] == out["metric"] and ref["value"] > 0
