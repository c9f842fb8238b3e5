"""CPU tier: the measurement contract of bench.py that can be checked without a GPU -- the own arm refuses to run
without a CUDA device (no CPU fallback anywhere in the product path), and the reference arm prints ONE JSON line with
the keys the driver reads, timing the reference's own Scene::trace (oracle/_ref when built here, else the oracle port)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_own_arm_fails_loudly_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0"], 300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not any(line.startswith("{") for line in r.stdout.splitlines())     # no bench line from a run that measured nothing


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("closest-hit Mrays/s")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "bunny20" in d["config"]["workload"] and d["steps"] == 1
