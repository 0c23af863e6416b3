"""CPU checks of bench.py's contract: the reference arm runs the compiled reference on host cores and prints
one JSON line with the keys the driver reads; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra, timeout=600):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, cwd=ROOT, timeout=timeout,
                          capture_output=True, text=True)


@pytest.mark.skipif(not O.have_reference(), reason="compiled reference (oracle/_ref) not built")
def test_reference_arm_prints_the_contract_line(tmp_path):
    small = {"XGM_BENCH_DOCS": "20000", "XGM_BENCH_VOCAB": "4000", "XGM_BENCH_REF_QUERIES": "48",
             "XGM_BENCH_REF_QUERIES_1T": "16", "XGM_REF_DB_DIR": str(tmp_path)}
    p = run_bench(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"], small)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["metric"].startswith("queries/sec") and line["unit"] == "queries/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert cb["single_thread"]["cores"] == 1 and cb["single_thread"]["value"] > 0
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit 0 without work
    q = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], dict(small, RANK="1", WORLD_SIZE="2"))
    assert q.returncode == 0 and not [l for l in q.stdout.splitlines() if l.startswith("{")]


def test_product_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = run_bench(["--gpus", "1", "--steps", "1", "--warmup", "1"], {})
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
