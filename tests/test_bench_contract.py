"""The reference arm of bench.py (CPU, no GPU needed): JSON contract of `bench.py --impl reference`, and that ranks other than 0 of a
torchrun launch exit without work or output."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None, args=("--steps", "1", "--warmup", "1")):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, timeout=900,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "clips/sec" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_only_rank_zero_works():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
