"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`: the
reference's algorithm on the host cores, oracle/ref_port.py) runs here end to end on a small graph and prints ONE JSON
line with the keys the driver reads; the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True, cwd=REPO,
                          timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--workload", "collab", "--scale", "0.01", "--steps", "2", "--warmup", "1", "--queries", "4",
             "--negs", "50")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scored links/sec" and d["unit"] == "links/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # links per step / time per step = the reported value
    links = 4 * 51
    assert abs(links / (d["ms_per_step"] * 1e-3) - d["value"]) / d["value"] < 1e-6


@pytest.mark.skipif(torch.cuda.is_available(), reason="the CUDA arm runs where there is a device")
def test_cuda_arm_refuses_to_run_without_a_device():
    r = _run("--workload", "collab", "--scale", "0.01", "--steps", "1", "--warmup", "0", "--queries", "2", "--negs", "10",
             "--no-cpu-baseline", "--no-sampler")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # no number without the GPU path


def test_reference_arm_under_torchrun_prints_one_line():
    """N > 1: rank 0 alone runs the reference arm and prints it, the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "collab", "--scale", "0.01", "--steps", "1", "--warmup", "1", "--queries", "4", "--negs", "50"],
                       capture_output=True, text=True, cwd=REPO, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
