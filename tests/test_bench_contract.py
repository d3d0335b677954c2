"""bench.py prints one JSON line with the contract's keys.  On CPU only the
reference arm can run (the product arm needs a GPU and must refuse without one)."""
import json
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, "bench.py", *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "200000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "matrices/s" and d["higher_is_better"] is True
    assert d["metric"] == "batched sym-solve matrices/sec" and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["steps"] == 2 and d["value"] > 0 and "workload" in d["config"]
    base = d["cpu_baseline"]
    # the reference's own code when it is importable (build container: /root/reference; GPU box: the
    # git-ignored baseline/_ref install), else the pinned port
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == d["value"] and "sample" in base
    assert d["e2e"] == {"value": d["value"], "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "3", "--batch", "1000")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
