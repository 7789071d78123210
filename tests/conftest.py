import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
BRICKS = ["neohook_brick", "a5_brick", "neohook_brick_analytical", "a5_brick_analytical"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build the CPU oracle (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port", "ref"], check=True)
    lib = os.path.join(ROOT, "fea-large_b200", "lib", "libfea_gpu.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "fea-large_b200"), "lib"], check=True)


def load_golden(name):
    from oracle.oracle import Model
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = Model(nodes=np.ascontiguousarray(z["nodes"]), conn=np.ascontiguousarray(z["conn"]),
              presc_node=np.ascontiguousarray(z["presc_node"]), presc_type=np.ascontiguousarray(z["presc_type"]),
              presc_vals=np.ascontiguousarray(z["presc_vals"]), model=int(z["model"]), lam=float(z["lam"]),
              mu=float(z["mu"]), gauss=int(z["gauss"]), load_increments=120,
              desired_tolerance=float(z["desired_tolerance"]), modified_newton=bool(z["modified_newton"]),
              max_newton=int(z["max_newton"]), solver_type=int(z["solver_type"]))
    return m, z


def load_brick_fine():
    """The reference's unstructured 22 934-tet model with the BC ids fixed (see make_golden.py)."""
    from oracle.oracle import Model
    z = np.load(os.path.join(GOLDEN, "brick_fine.npz"))
    m = Model(nodes=np.ascontiguousarray(z["nodes_nano"] / 1e9), conn=np.ascontiguousarray(z["conn"]),
              presc_node=np.ascontiguousarray(z["presc_node"]), presc_type=np.ascontiguousarray(z["presc_type"]),
              presc_vals=np.ascontiguousarray(z["presc_vals"]), model=int(z["model"]), lam=float(z["lam"]),
              mu=float(z["mu"]), gauss=int(z["gauss"]), load_increments=1, desired_tolerance=1e-6,
              modified_newton=True, max_newton=1, solver_type=0)
    x = m.nodes.copy()
    x[:, 1] = 1.0 + (x[:, 1] - 1.0) * 1.01          # the 1 % stretched state the golden was taken on
    return m, z, x


@pytest.fixture(params=BRICKS)
def brick(request):
    m, z = load_golden(request.param)
    return request.param, m, z


def csr_mv(rp, ci, v, x):
    y = np.zeros(len(rp) - 1)
    np.add.at(y, np.repeat(np.arange(len(rp) - 1), np.diff(rp)), v * x[ci])
    return y


def block_model(n, model=1, bc_style=0, dy=0.0, box=(1.0, 1.0, 1.0), y0=0.0):
    """Kuhn block as an oracle Model (uses the product's host mesh generator)."""
    import fea_gpu as fg
    from oracle.oracle import Model
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    mb = fg.mesh_block(nx, ny, nz, box[0], box[1], box[2], y0, bc_style, dy)
    return Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=mb["presc_node"], presc_type=mb["presc_type"],
                 presc_vals=mb["presc_vals"], model=model, lam=100.0, mu=100.0, gauss=5)


def write_sexp(path, m, load_increments=2, shuffle_keys=False):
    """Write an oracle Model as a task file in the reference's .sexp grammar
    (utilities/exporter.py:442-502 is the reference's writer; ids are 0-based)."""
    name = "COMPRESSIBLE_NEOHOOKEAN" if m.model == 1 else "A5"
    solver = {0: "CG", 1: "PCG_ILU", 2: "CHOLESKY"}[m.solver_type]
    with open(path, "w") as f:
        f.write(";; -*- Mode: lisp; -*-\n(task\n")
        f.write(f" (model :name {name}\n        (model-parameters :mu {m.mu:g} :lambda {m.lam:g}))\n")
        if shuffle_keys:   # key order is free; symbols are case-insensitive; ';' comments
            f.write(f" (solution :max-newton-count {m.max_newton} :modified-newton {'Yes' if m.modified_newton else 'no'} ; comment\n"
                    f"   :load-increments-count {load_increments} :task-type cartesian3d :desired-tolerance {float(m.desired_tolerance)!r}\n")
        else:
            f.write(f" (solution :desired-tolerance {float(m.desired_tolerance)!r} :task-type CARTESIAN3D "
                    f":load-increments-count {load_increments} :modified-newton {'yes' if m.modified_newton else 'no'} "
                    f":max-newton-count {m.max_newton}\n")
        f.write(f"   (element-type :gauss-nodes-count {m.gauss} :name TETRAHEDRA10 :nodes-count 10)\n")
        f.write(f"   (slae-solver :type {solver} :tolerance {float(m.solver_tolerance)!r} :max-iterations {m.solver_max_iter})\n")
        f.write(f"   (line-search :max {int(m.extra.get('linesearch', 0))})\n   (arc-length :max 0))\n (input-data\n  (geometry\n   (nodes\n")
        for x in m.nodes:
            f.write(f"    ({float(x[0])!r} {float(x[1])!r} {float(x[2])!r})\n")
        f.write("   )\n   (elements\n")
        for c in m.conn:
            f.write("    (" + " ".join(str(int(v)) for v in c) + ")\n")
        f.write("   ))\n  (boundary-conditions\n   (prescribed-displacements\n")
        for n, t, v in zip(m.presc_node, m.presc_type, m.presc_vals):
            f.write(f"    (presc-node :y {float(v[1])!r} :x {float(v[0])!r} :z {float(v[2])!r} :type {int(t)} :node-id {int(n)})\n")
        f.write("   ))))\n")
