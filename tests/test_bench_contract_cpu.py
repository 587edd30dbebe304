"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference TU on the host cores, the one place
besides tests/ and smoke() that may execute oracle/) prints ONE JSON line with the contract's keys, and the
product arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpicsp_ref.so")


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


@pytest.mark.skipif(not os.path.isfile(REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--cells", "64", "--cpu-particles", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["value"] > 1e5 and d["ms_per_step"] > 0


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0", "--cells", "64", "--particles", "2000", "--no-e2e", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")], "no bench line may be printed without a GPU"
