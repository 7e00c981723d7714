"""bench.py's reference arm (the CPU restatement timed on the host cores) prints the contract's
JSON line; runs here without a GPU on a tiny sample.  The GPU arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, timeout=300, env=e)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--n-db", "20000",
                  "--queries", "64", "--k", "10", "--cpu-sample-rows", "20000",
                  "--cpu-sample-queries", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "cpu_baseline", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # the reference's own CPU path (scikit-learn brute force) is reported beside the port
    sk = line["cpu_baseline_sklearn"]
    assert sk["kind"] == "reference" and sk["value"] > 0 and "NearestNeighbors" in sk["sample"]


def test_reference_arm_other_ranks_print_nothing():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1", "--warmup", "0", "--n-db", "1000", "--queries", "8", "--k", "4")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches N > 1 through torch.distributed.run: rank 0 alone works and prints,
    the other ranks exit 0 without work."""
    e = dict(os.environ)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--n-db", "20000",
         "--queries", "32", "--k", "10", "--cpu-sample-rows", "20000", "--cpu-sample-queries", "32"],
        capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
