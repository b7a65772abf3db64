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


def test_image_error_gates_separate_an_unbiased_from_a_biased_image():
    """bench.py's per-scene image error: for an estimator with the reference's own statistics the 8x8-block RMSE
    sits at the value the reference's two half-buffers predict (ratio ~ 1) and both gates pass; a 3 % brightness bias or extra
    block-scale structure fails them."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", BENCH)
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(1)
    h, w = 256, 320
    truth = 0.2 + 0.8 * rng.random((h // 8, w // 8, 3)).repeat(8, axis=0).repeat(8, axis=1)      # block-constant image
    sigma1 = 0.7                                                                                      # noise of one sample per pixel

    def mean_image(spp):
        return (truth + rng.normal(0, sigma1 / np.sqrt(spp), truth.shape)).astype(np.float32)
    n, m_half = 64, 8
    ha, hb = mean_image(m_half), mean_image(m_half)
    good = bench.image_error(mean_image(n), n, ha, hb, m_half)
    assert 0.9 < good["block8_rmse_over_expected"] < 1.1
    assert good["gate_mean_0p5pct_or_3x_cpu_noise"] and good["gate_block_rmse_3x_noise"]
    assert good["gpu_spp"] == n and good["cpu_spp"] == 2 * m_half and good["relmse"] > 0
    bright = bench.image_error(mean_image(n) * np.float32(1.03), n, ha, hb, m_half)
    assert not bright["gate_mean_0p5pct_or_3x_cpu_noise"]
    blotchy = mean_image(n) + 0.2 * rng.normal(0, 1, (h // 8, w // 8, 3)).repeat(8, axis=0).repeat(8, axis=1).astype(np.float32)
    bad = bench.image_error(blotchy, n, ha, hb, m_half)
    assert not bad["gate_block_rmse_3x_noise"] and bad["block8_rmse_over_expected"] > 3
