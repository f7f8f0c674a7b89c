"""CPU: bench.py's output contract — the reference arm prints exactly one JSON line on stdout with the agreed keys,
and the GPU arm refuses to run (no CPU fallback) when no CUDA device is present."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "1", "--ref-reads", "120"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads/sec mapped" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly(built):
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_human_size_reference_arm_says_why_it_needs_the_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == "" and "index table from the GPU builder" in r.stderr


def test_batch_size_follows_the_time_budget(monkeypatch):
    """The human-size batch shrinks when the steps asked for would not fit the run's time budget (the driver runs
    --steps 20 --warmup 5); explicit sizes win; the yeast-size config has no budget rule."""
    import importlib.util
    import types
    monkeypatch.delenv("RH_BENCH_READS_HUMAN", raising=False)
    monkeypatch.delenv("RH_BENCH_BUDGET_S", raising=False)
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    cfgs = b._configs()
    A = types.SimpleNamespace
    assert b._reads_per_step(cfgs[2], A(reads=0, steps=2, warmup=3)) == 24000
    r20 = b._reads_per_step(cfgs[2], A(reads=0, steps=20, warmup=5))
    assert 8000 <= r20 <= 14000
    total_steps = 5 + 20 + min(20, b.E2E_STEPS_MAX) + 2
    assert total_steps * r20 / cfgs[2]["rate_est"] <= 540
    assert b._reads_per_step(cfgs[3], A(reads=0, steps=20, warmup=5)) < r20          # the R10 workload is slower per read
    assert b._reads_per_step(cfgs[2], A(reads=0, steps=500, warmup=5)) == 4000       # floor
    assert b._reads_per_step(cfgs[2], A(reads=1234, steps=20, warmup=5)) == 1234
    assert b._reads_per_step(cfgs[1], A(reads=0, steps=20, warmup=5)) == 100000
    monkeypatch.setenv("RH_BENCH_BUDGET_S", "2000")
    assert b._reads_per_step(cfgs[2], A(reads=0, steps=20, warmup=5)) == 24000
