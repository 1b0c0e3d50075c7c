"""CPU: bench.py's behaviour without a GPU.  Our arm refuses to run (no CPU fallback on the product path); the reference arm
falls back to the CPU oracle port on a bounded sample and prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="describes the behaviour of a machine without a GPU")


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_our_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "ours" and "no CPU fallback" in line["error"]


def test_reference_arm_times_the_cpu_port_on_a_bounded_sample():
    r = _run("--impl", "reference", "--size", "mid", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_second_foveated_1080p_6M" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "3 frames" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("fov_300k_800x600")


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs the CPU port and prints the line; the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "mid",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
