"""`bench.py --impl reference` without a GPU (CPU suite): the reference arm must run on the host alone -- the unmodified
reference through oracle/_ref -- and print ONE JSON line with the contract's keys, for the default workload (configs[3]) at a
reduced size; under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_ctypes as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")

KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def run(env_extra, *flags):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags], capture_output=True,
                       text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_default_workload_small():
    lines = run({}, "--map-points", "40000", "--queries", "20000", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1  # the reference prints from C (`Multi thread started`): stdout must stay one JSON line
    out = json.loads(lines[0])
    for k in KEYS:
        assert k in out, k
    assert out["impl"] == "reference" and out["config"]["workload"].startswith("configs[3]") and out["scaling"] == "strong"
    assert out["value"] > 0 and out["higher_is_better"] is True and out["unit"] == "queries/s"
    assert out["cpu_baseline"]["kind"] == "reference" and out["cpu_baseline"]["cores"] >= 1
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--map-points", "40000", "--queries", "20000") == []
