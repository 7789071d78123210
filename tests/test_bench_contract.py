"""bench.py prints ONE JSON line with the keys the driver reads -- checked on a tiny workload.
CPU: the reference arm (the reference's compiled element code, or the port).  GPU: the own arm."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=900)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines            # exactly one line on stdout, whatever libraries print
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "3")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "element_assemblies_per_sec" and d["unit"] == "elements/s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_own_arm_line_small_workload():
    d = run_bench("--n", "10", "--steps", "3", "--warmup", "3", "--newton-iters", "1", "--cpu-baseline-sample", "3")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak"
    # per step: x += u, element pass, two gathers (Dirichlet cancellation folded in)
    assert d["value"] > 1e6 and d["gpu_launches"] >= 3 * 4 and d["bad_points"] == 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] == 24 * 21 ** 3 == e["d2h_bytes_per_step"]
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0
    n = d["newton"]
    assert n["pcg_relres"] <= 1e-13 and n["newton_iters_per_sec"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
