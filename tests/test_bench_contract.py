"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`) prints exactly one
JSON line with the keys the driver reads, and the GPU arm refuses to run without a device (no CPU path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    res = _run("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lanczos_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("c1:")
    cb = d["cpu_baseline"]
    # "reference": the unmodified reference classes vendored into oracle/_ref by oracle/build_ref.py (present
    # wherever __graft_entry__.build() ran with /root/reference in reach); "port": the oracle's restatement
    from oracle import build_ref
    assert cb["kind"] == ("reference" if build_ref.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    res = _run("--impl", "reference", "--workload", "c1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA device present")
    res = _run("--workload", "c1", "--steps", "2")
    assert res.returncode != 0
    assert "no CPU path" in (res.stderr + res.stdout)
