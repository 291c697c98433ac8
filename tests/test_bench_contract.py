"""bench.py's output contract on the CPU side: the reference arm prints exactly ONE JSON line on stdout (whatever
libraries print goes to stderr), with the keys the driver reads; under a multi-rank launch only rank 0 prints; the
native arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "grid100"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sim link-steps/s" and d["unit"] == "link-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "grid100" and d["config"]["links"] == 39600 and d["config"]["dual_edges"] == 157208


def test_reference_arm_other_ranks_print_nothing():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "grid100"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_native_arm_refuses_to_run_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
