"""bench.py on a box without a GPU: the reference arm (the CPU port on the host cores) runs and prints ONE JSON line
with the contract's keys; the CUDA arm refuses loudly (there is no CPU fallback of the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = run("--impl", "reference", "--voices", "64", "--seconds", "0.25", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rendered voice-samples/sec" and j["unit"] == "voice-samples/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_cuda_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = run("--steps", "1", "--warmup", "1", "--voices", "64", "--seconds", "0.1", "--no-extras")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
