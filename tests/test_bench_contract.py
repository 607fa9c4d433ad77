"""bench.py's contract on a machine without a GPU: the reference arm (the reference's CPU
implementation of the path = the oracle port, timed on the host cores) prints exactly one JSON
line with the agreed keys, and the product arm refuses to run (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "3"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "RoIs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1
    assert d["config"]["workload"].startswith("cfg2")
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_product_arm_has_no_cpu_fallback():
    p = _run(["--steps", "1", "--warmup", "3", "--no-cfg3", "--no-cpu-baseline"], timeout=120)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
