"""CPU: the C-ABI library loads without a GPU and exports every symbol include/fea_gpu.h
declares; argument validation works; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fea_gpu as fg
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "fea_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fea_(?:gpu|plan|mesh)_\w+)\s*\(", text)))


def test_exports_every_declared_symbol():
    L = fg.lib()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fea_gpu.h but not exported"
    assert sorted(fg.SYMBOLS) == names


def test_launch_count_starts_at_zero_on_cpu():
    assert fg.launch_count() >= 0


def test_create_validates_arguments():
    mb = fg.mesh_block(1, 1, 1)
    with pytest.raises(fg.FeaGpuError) as e:
        fg.FeaGpu(mb["nodes"], mb["conn"], model=9, lam=1, mu=1)
    assert e.value.code == fg.ERR_ARG
    with pytest.raises(fg.FeaGpuError) as e:
        fg.FeaGpu(mb["nodes"], mb["conn"], model=0, lam=1, mu=1, n_gauss=3)   # fea_solver.c:1503
    assert e.value.code == fg.ERR_ARG


def test_no_cpu_fallback_without_device():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import fea_gpu as fg\n"
            "mb = fg.mesh_block(1,1,1)\n"
            "try:\n    fg.FeaGpu(mb['nodes'], mb['conn'], 0, 100., 100.)\n    print('CREATED')\n"
            "except fg.FeaGpuError as e:\n    print('ERR', e.code)\n") % os.path.join(ROOT, "fea-large_b200", "python")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout
    assert out.strip() == f"ERR {fg.ERR_CUDA}"


def test_product_never_links_the_oracle():
    so = os.path.join(ROOT, "fea-large_b200", "lib", "libfea_gpu.so")
    import subprocess
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "orc_" not in syms and "ref_" not in syms.replace("fea_gpu_", "")
    for root, _, files in os.walk(os.path.join(ROOT, "fea-large_b200")):
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".cpp", ".hpp", ".py")):
                text = open(os.path.join(root, f)).read()
                assert "oracle_fea" not in text and "libfea_ref" not in text and "libfea_oracle" not in text, f
