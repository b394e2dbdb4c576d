"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.gpu
def test_bench_line_has_the_contract_keys(cuda_device):
    d = _run("--gpus", "1", "--steps", "2", "--warmup", "3", "--times", "1", "--e2e-times", "1",
             "--cpu-bl", "16", "--cpu-times", "4")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "hera350" in d["config"]["workload"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.05 < r["frac"] < 1.2 and d["gpu_launches"] == 4 and d["value"] > 1e9
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0 and "sample" in c
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # host buffers + PCIe can only be slower than HBM-resident
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # round 2: every BASELINE.json configuration in the same line, the fixed job, the H2D ceiling
    assert set(r["modes"]) >= {"hera350_complex128", "hera350_nuv2_complex64", "ska512_complex64", "ska512_complex128",
                               "paper64_203ch_complex64"}
    assert r["modes"]["paper64_203ch_complex64"]["nfreq"] == 203 and r["modes"]["ska512_complex64"]["nfreq"] == 2048
    assert all(m["fused"] and 0.01 < m["frac"] < 1.2 for m in r["modes"].values())
    assert set(r["legs"]) == {"transform+pair_sums", "transform+pair_sums+thermal", "transform+pair_sums+noise"}
    fj = d["fixed_job"]
    assert fj["times"] == 1000 and fj["scaling"] == "strong" and fj["value"] > 1e9 and "clocks" in fj
    assert 0 < e["frac_of_h2d_ceiling"] < 1.3 and e["d2h_bytes_per_job"] >= e["d2h_bytes_per_step"]
    assert d["multi_gpu_check"] is None  # one GPU


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-bl", "8", "--cpu-times", "2")
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "pair-spectra/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
