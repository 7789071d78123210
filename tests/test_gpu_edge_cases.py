"""GPU: corner cases of the reference's semantics, CUDA path vs the CPU oracle."""
import numpy as np
import pytest

import fea_gpu as fg
from conftest import block_model, load_golden
from oracle.oracle import Model, PortOracle
from test_gpu_parity import RTOL_ELEM, RTOL_SOLVE, deformed, make_gpu, relmax

pytestmark = pytest.mark.gpu


def both(m, x=None, lam_inc=1.0):
    g, o = make_gpu(m), PortOracle(m)
    for s in (g, o):
        if x is not None:
            s.set_nodes(x)
        s.apply_increment(lam_inc); s.update_state(); s.assemble_stiffness(); s.assemble_residual()
    return g, o


def test_duplicate_free_and_unknown_bc_entries():
    """solver_apply_bc_general (fea_solver.c:1205-1242) walks the LIST: a node listed twice is moved
    twice, type 0 (FREE) and values outside the enum select no DOF."""
    m = block_model((2, 2, 2), model=1, dy=0.01)
    extra_n = np.array([m.presc_node[3], m.presc_node[3], m.presc_node[5], 40], np.int32)
    extra_t = np.array([2, 5, 0, 9], np.int32)
    extra_v = np.array([[0.3, 0.02, 0.1], [0.01, 9.0, -0.02], [1, 1, 1], [1, 1, 1]], float)
    m = Model(nodes=m.nodes, conn=m.conn, presc_node=np.concatenate([m.presc_node, extra_n]),
              presc_type=np.concatenate([m.presc_type, extra_t]),
              presc_vals=np.concatenate([m.presc_vals, extra_v]), model=1)
    g, o = both(m)
    assert np.array_equal(g.get_nodes(), o.get_nodes())                 # increments summed per DOF
    for s in (g, o):
        s.apply_bc(0.0)
    assert relmax(g.get_csr()[3], o.get_csr()[2]) < RTOL_ELEM
    assert relmax(g.get_forces(), o.get_forces()) < 10 * RTOL_ELEM
    g.solve(1e-14, 5000); o.solve_slae()
    assert relmax(g.get_solution(), o.get_solution()) < RTOL_SOLVE


def test_no_prescribed_nodes_at_all():
    m = block_model((2, 2, 2), model=0)
    m = Model(nodes=m.nodes, conn=m.conn, presc_node=np.zeros(0, np.int32), presc_type=np.zeros(0, np.int32),
              presc_vals=np.zeros((0, 3)), model=0)
    g, o = both(m, deformed(m, 8, 0.01))
    v0 = g.get_csr()[3].copy()
    g.apply_bc(0.0); o.apply_bc(0.0)
    assert np.array_equal(g.get_csr()[3], v0) and relmax(v0, o.get_csr()[2]) < RTOL_ELEM


def test_single_element_mesh():
    m = block_model((1, 1, 1), model=1)
    m = Model(nodes=m.nodes, conn=np.ascontiguousarray(m.conn[:1]), presc_node=np.array([int(m.conn[0, 0])], np.int32),
              presc_type=np.array([7], np.int32), presc_vals=np.zeros((1, 3)), model=1)
    g, o = both(m, deformed(m, 3, 0.02))
    rows, rp, ci, v = g.get_csr()
    rpo, cio, vo = o.get_csr()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio) and relmax(v, vo) < RTOL_ELEM
    assert len(v) == 900                    # 10 coupled nodes; the 17 unused nodes have empty rows, as in the reference
    assert relmax(g.get_forces(), o.get_forces()) < RTOL_ELEM


def test_four_point_rule_on_shipped_brick():
    m, _ = load_golden("neohook_brick")
    m.gauss = 4                              # gauss_nodes4_tetr10: 8-digit literals (fea_solver.c:32-48)
    g, o = both(m, deformed(m, 6))
    assert relmax(g.get_csr()[3], o.get_csr()[2]) < RTOL_ELEM
    assert relmax(g.get_forces(), o.get_forces()) < RTOL_ELEM
    Fg, Sg = g.get_state(); Fo, So = o.get_state()
    assert Fg.shape == (346, 4, 3, 3) and relmax(Sg, So) < RTOL_ELEM


def test_zero_right_hand_side_and_restart_vectors():
    m, _ = load_golden("a5_brick")
    g = make_gpu(m)
    g.assemble_all(True); g.apply_bc(0.0)                 # undeformed: R is rounding noise (F = I to 1e-16)
    assert np.abs(g.get_forces()).max() < 1e-12
    it, rr, ok = g.solve(1e-14, 5000, accept_stall=True)   # a noise right-hand side may end on the rounding floor
    assert ok and np.abs(g.get_solution()).max() < 1e-13
    g.set_forces(np.zeros(m.n_dof))                       # an exactly zero right-hand side: nothing to do
    it, rr, ok = g.solve(1e-14, 100)
    assert ok and it == 0 and np.all(g.get_solution() == 0.0)
    g.set_nodes(m.nodes)
    g.apply_increment(1.0); g.assemble_all(True); g.apply_bc(0.0)
    it0, _, _ = g.solve(1e-13, 5000, fg.X0_ZERO); u0 = g.get_solution()
    it1, _, _ = g.solve(1e-13, 5000, fg.X0_RHS); u1 = g.get_solution()
    assert relmax(u1, u0) < 1e-9
    itabs, rr, ok = g.solve(1e-9, 5000, fg.ABS_TOL)        # absolute tolerance: ||r|| <= 1e-9
    R = g.get_forces()
    assert ok and np.linalg.norm(g.spmv(g.get_solution()) - R) <= 2e-9


def test_saved_stiffness_round_trip():
    """sp_matrix_copy at fea_solver.c:179 / :194-195 (modified Newton)."""
    m, _ = load_golden("neohook_brick")
    g = make_gpu(m)
    g.apply_increment(1.0); g.assemble_all(True); g.save_stiffness()
    k0 = g.get_csr()[3].copy()
    g.apply_bc(0.0)
    assert not np.array_equal(g.get_csr()[3], k0)
    g.restore_stiffness()
    assert np.array_equal(g.get_csr()[3], k0)
    h = make_gpu(m)
    with pytest.raises(fg.FeaGpuError):
        h.restore_stiffness()                               # nothing saved yet
