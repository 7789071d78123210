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
