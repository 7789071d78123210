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


def test_reference_arm_never_loads_the_product_library():
    """The reference arm is the reference's CPU code only: the repo's CUDA library must not be mapped."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--ref-sample', '2'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libfea_gpu' not in maps, 'product library mapped'; "
            "assert 'fea_gpu' not in sys.modules") % os.path.join(ROOT, "bench.py")
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]


def test_window_parity_logic_on_a_cpu_stand_in():
    """bench.py's parity block on a small mesh with the GPU context replaced by an oracle of the FULL mesh:
    the window's F, sigma, K_e, interior rows of R and K probes must agree to rounding, i.e. the window
    extraction, the id-keyed bench state and the 'interior rows are complete' argument are right."""
    import types
    import numpy as np
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    # bench.py redirects fd 1 at import: keep pytest's capture intact
    saved = os.dup(1)
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    from oracle.kuhn import kuhn_block
    from oracle.oracle import Model, PortOracle
    n, world = 5, 2
    mb = kuhn_block(n, n * world, n, float(n), float(n * world), float(n), 0.0, 1, 0.01)
    m = Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=np.zeros(0, np.int32), presc_type=np.zeros(0, np.int32),
              presc_vals=np.zeros((0, 3)), model=0, lam=100.0, mu=100.0, gauss=5)
    full = PortOracle(m)
    full.set_nodes(bench.deformed_state(m.nodes, 0.5, model=0))

    class Stand:
        def assemble_all(self, with_k):
            full.update_state(); full.assemble_stiffness(); full.assemble_residual()
            self.csr = full.get_csr()
        def get_state_elems(self, elems):
            F, S = full.get_state()
            return F[elems], S[elems], np.ones(len(elems), bool)
        def element_matrix(self, e): return full.element_matrix(e)
        def get_forces(self): return full.get_forces()
        def spmv(self, x):
            rp, ci, v = self.csr
            y = np.zeros(len(rp) - 1)
            np.add.at(y, np.repeat(np.arange(len(rp) - 1), np.diff(rp)), v * x[ci])
            return y

    args = types.SimpleNamespace(n=n, parity_sample=4, parity_elems=16, model=0)
    err, info = bench.window_parity(Stand(), args, world, 0, len(m.nodes))
    assert info["window_elements"] == 6 * 4 ** 3 and info["interior_rows"] == 3 * 7 ** 3
    assert info["window_origin_cubes"] == [0, 3, 0]            # straddles the rank interface at cube y = n
    assert max(err.values()) < 1e-13, err


@pytest.mark.gpu
def test_own_arm_line_small_workload():
    d = run_bench("--n", "10", "--steps", "3", "--warmup", "3", "--newton-iters", "1", "--cpu-baseline-sample", "3",
                  "--cpu-newton-sample", "3", "--c4-n", "8", "--parity-sample", "4")
    p = d["parity"]
    assert p["ok"] and p["max_rel_elem"] <= 1e-11 and p["bench_mesh_window"]["K_e_compared_all_ranks"] > 0
    assert d["strong_c4"]["pcg_exit"] == 1 and d["strong_c4"]["elements"] == 6 * 8 ** 3 and d["comm"]["halo_ms"] < 1e-3
    assert d["cpu_baseline"]["newton"]["pcg_iters"] > 0 and d["roofline_fp64"]["dmma_m8n8k4_tflops_this_run"] > 0
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
