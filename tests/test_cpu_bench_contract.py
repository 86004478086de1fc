"""bench.py's reference arm on the host cores (no GPU needed): the JSON line the driver parses carries the contract's keys, the
arm runs the CPU port of the path (`cpu_baseline.kind`), and -- where oracle/_ref was built -- the compiled reference's own
PnPsolver.cpp beside it.  Without a GPU the product arm must refuse to run rather than fall back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], {"RSAC_BENCH_REF_SWEEPS": "1", "RSAC_BENCH_COMPILED_REF_N": "64"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "candidates/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "cfg4" and d["config"]["matches"] == 500 and d["config"]["hypotheses"] == 300
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_solvers.so")):
        c = d["compiled_reference"]
        assert c["kind"] == "reference" and c["value"] > 0 and c["candidates_ok"] > 32      # most cfg4 candidates verify


def test_product_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
