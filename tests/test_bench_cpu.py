"""bench.py on a box without a GPU: the reference arm (the C port of the reference's CPU algorithm) runs and prints the contract's
JSON line; the product arm refuses to run (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, cwd=ROOT, env=e, timeout=900)


def test_reference_arm_prints_the_contract_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "proofs/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["metric"].startswith("R1CS proofs/sec") and line["vs_baseline"] is None
    assert "n=18176, N=32768, m=69, q=42369" in line["config"]["workload"]  # the depth-32 membership circuit
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_the_device():
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = run("--steps", "1", "--warmup", "0", "--no-extras", "--no-cpu-baseline")
    assert r.returncode != 0  # no CPU path: the library reports BP_ERR_NO_DEVICE and the bench stops
    assert r.stdout.strip() == "" or "value" not in r.stdout
