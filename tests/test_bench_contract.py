"""The bench.py contract (ONE JSON line on stdout, the keys the driver reads) -- CPU arm here, GPU arm under -m gpu."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}: {r.stdout[:500]}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--config", "mini", "--steps", "4", "--warmup", "3"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "IPM iterations/sec (KKT factor+solve)" and d["unit"] == "iter/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] == "port" and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["ipm"]["status"] == "Trm_Optimal" and d["ipm"]["rel_gap"] < 1e-7
    assert "NOT Tulip/CHOLMOD" in cb["label"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "mini"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "mini"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--config", "mini", "--steps", "4", "--warmup", "3"])
    assert BASE_KEYS <= set(d)
    for k in ("clocks", "gpu_launches", "roofline", "cpu_baseline", "steps_real", "ipm", "roofline_solve", "phases_one_step"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["ipm"]["status"] == "Trm_Optimal"
    assert d["ipm_device_resident"]["status"] == "Trm_Optimal"
