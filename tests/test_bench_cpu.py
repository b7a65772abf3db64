"""bench.py's command-line contract, as far as it can be checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_own_arm_without_a_gpu_fails_loudly():
    """No CUDA device -> non-zero exit and a message; there is no CPU fallback to time instead."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = run([])
    assert out.returncode != 0
    assert "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr
    assert out.stdout.strip() == ""


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    out = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-spp", "1"])
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None


def test_reference_arm_on_other_ranks_exits_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work."""
    out = run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert out.returncode == 0
    assert out.stdout.strip() == ""
