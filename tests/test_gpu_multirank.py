"""GPU, needs >= 2 devices (skipped on a one-GPU box): the NCCL path -- halo exchange of
coordinates and of the CG direction, all-reduced dot products, all-gathered read-back --
against the single-rank device result and the CPU oracle."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_multirank_newton_step_matches_single_rank(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "multirank_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert "MULTIRANK_RESULT PASS" in run.stdout, run.stdout[-3000:] + run.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_single_process_multi_gpu_handle(world):
    """fea_gpu_create_multi: ONE process drives `world` GPUs (one context + host thread per GPU inside the
    library).  The handle must behave like a single context: same R, u, <R,u>, x, sigma, K_e as one GPU."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    code = f"""
import sys, numpy as np
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'fea-large_b200', 'python')!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import fea_gpu as fg
from multirank_worker import small_bar, relmax
m = small_bar({world})
x0 = m.nodes + 0.004 * np.random.default_rng(3).standard_normal(m.nodes.shape)
res = []
for kw in (dict(n_gpus={world}), dict(device=0)):
    g = fg.FeaGpu(m.nodes, m.conn, m.model, m.lam, m.mu, 5, m.presc_node, m.presc_type, m.presc_vals, **kw)
    g.set_nodes(x0); g.apply_increment(1.0); g.assemble_all(True)
    R0 = g.get_forces(); ke = g.element_matrix(7); F, S, found = g.get_state_elems(np.arange(len(m.conn), dtype=np.int32))
    g.apply_bc(0.0); it, rr, ok = g.solve(1e-13, 20000); tol = g.dot_R_u(); u = g.get_solution()
    g.update_nodes(); g.assemble_all(True)
    x1 = g.get_nodes(); Fa, Sa = g.get_state(); bad = g.bad_points(); cnt = g.counts()
    res.append(dict(R0=R0, ke=ke, S=S, found=found, u=u, tol=tol, x1=x1, Sa=Sa, it=it, ok=ok, bad=bad, cnt=cnt))
    g.close()
a, b = res
assert a['cnt']['owned_nodes'] == b['cnt']['owned_nodes'] == len(m.nodes) and a['found'].all() and a['ok'] and b['ok']
e = dict(R0=relmax(a['R0'], b['R0']), ke=relmax(a['ke'], b['ke']), S=relmax(a['S'], b['S']), u=relmax(a['u'], b['u']),
         tol=abs(a['tol'] - b['tol']) / abs(b['tol']), x1=relmax(a['x1'] - m.nodes, b['x1'] - m.nodes), Sa=relmax(a['Sa'], b['Sa']))
print('MULTI_HANDLE', a['it'], b['it'], e, flush=True)
assert e['R0'] < 1e-12 and e['ke'] < 1e-12 and e['S'] < 1e-12 and e['u'] < 1e-9 and e['tol'] < 1e-9 and e['x1'] < 1e-9 and e['Sa'] < 1e-9
assert a['bad'] == 0 and b['bad'] == 0
print('MULTI_HANDLE PASS', flush=True)
"""
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert "MULTI_HANDLE PASS" in run.stdout, run.stdout[-3000:] + run.stderr[-3000:]


def test_feasolver_binary_on_two_gpus_writes_the_same_file(tmp_path):
    """FEA_GPU_COUNT=2: the drop-in binary (one process, as the reference's do_main) drives two GPUs and
    exports what it exports on one."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import load_golden, write_sexp
    m, _ = load_golden("neohook_brick")
    path = str(tmp_path / "two.sexp")
    write_sexp(path, m, load_increments=2)
    binary = os.path.join(ROOT, "fea-large_b200", "bin", "feasolver_b200")
    outs = []
    for n in (1, 2):
        run = subprocess.run([binary, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=900,
                             env=dict(os.environ, FEA_GPU_COUNT=str(n)))
        assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
        outs.append(open(str(tmp_path / "two.msh")).read().splitlines())
    assert len(outs[0]) == len(outs[1])
    worst = 0.0
    for a, b in zip(*outs):
        if a != b:
            na, nb = [float(v) for v in a.split()], [float(v) for v in b.split()]
            worst = max(worst, max(abs(p - q) for p, q in zip(na, nb)))
    assert worst <= 2e-6, worst
