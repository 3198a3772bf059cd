"""bench.py's reference arm (the only arm that runs without a GPU): the JSON line carries the keys the driver's
contract names, `config` is arm-independent, and the numbers are internally consistent.  The GPU arm is exercised by
the driver on a B200; here we also check that it refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

import oracle_ffi as of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    return out


@pytest.mark.skipif(not of.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line_has_the_contract_keys():
    out = run_bench("--impl", "reference", "--kf", "8", "--pts", "300", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "lm_iterations_per_s" and d["unit"] == "LM it/s"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None                       # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert cb["final_cost"] < cb["initial_cost"]


def test_gpu_arm_refuses_to_run_without_a_device():
    import pba_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("only meaningful on a box without a GPU")
    out = run_bench("--kf", "8", "--pts", "300", "--steps", "1", "--warmup", "1")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
