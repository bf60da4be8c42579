"""bench.py keeps the driver's JSON contract: the reference arm (the only arm that runs without a GPU) prints ONE line
with the agreed keys for every workload; CPU only, bounded to a few seconds per workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


@pytest.mark.parametrize("workload,size", [("md3d", 0), ("md2d", 0), ("st1d", 200001), ("st1d_burgers", 200001), ("e1d", 200001)])
def test_reference_arm_json_line(workload, size):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
           "--warmup", "0", "--ref-budget", "0.5", "--ref-replicas", "2" if workload in ("md3d", "st1d") else "0"] \
        + (["--size", str(size)] if size else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == "updates/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["single_thread"] > 0
    if workload in ("md3d", "st1d"):      # all the host threads the single-threaded reference can use: independent replicas
        allc = d["cpu_baseline"]["all_cores"]
        assert allc["cores"] == 2 and allc["value"] > 0 and allc["unit"] == d["unit"]
        assert d["cpu_baseline"]["cores"] == 2 and d["value"] == allc["value"] == d["cpu_baseline"]["value"]
    else:
        assert "all_cores" not in d["cpu_baseline"] and d["cpu_baseline"]["cores"] == 1
        assert d["value"] == d["cpu_baseline"]["single_thread"]
    assert "reference_sample_note" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
