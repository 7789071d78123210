"""CPU, only where oracle/_ref exists (built from /root/reference): the plain-C oracle
must agree BITWISE with the reference's own compiled objects."""
import numpy as np
import pytest

from oracle.oracle import PortOracle, RefOracle, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def test_reference_self_test_passes():
    assert RefOracle.lib().ref_do_tests() == 1        # solver-large/tests.c, run by main() at startup


def test_port_equals_reference_everywhere(brick):
    name, m, _ = brick
    r, p = RefOracle(m), PortOracle(m)
    rng = np.random.default_rng(7)
    x = m.nodes + 0.01 * rng.standard_normal(m.nodes.shape)   # a non-trivial deformed state
    for o in (r, p):
        o.set_nodes(x)
        o.apply_increment(1.0)
        o.update_state()
        o.assemble_stiffness()
        o.assemble_residual()
    for a, b in zip(r.get_state() + r.get_gradients() + r.get_csr() + (r.get_forces(),),
                    p.get_state() + p.get_gradients() + p.get_csr() + (p.get_forces(),)):
        assert np.array_equal(a, b)
    for e in (0, 100, 345):
        for part in (0, 1, 2):
            assert np.array_equal(r.element_matrix(e, part), p.element_matrix(e, part))
    for o in (r, p):
        o.apply_bc(0.0)
        o.solve_slae()
    assert np.array_equal(r.get_csr()[2], p.get_csr()[2])
    assert np.array_equal(r.get_solution(), p.get_solution())


def test_newton_trajectory_full_newton():
    from conftest import load_golden
    m, _ = load_golden("neohook_brick")
    rhs, sol, tol = RefOracle.run_solve(m, load_increments=1, modified_newton=False, desired_tol=1e-12)
    p = PortOracle(m)
    done, tu, tt = p.newton_solve(1, 1e-12, False, m.max_newton)
    assert done == 1 and np.array_equal(tu, sol) and np.array_equal(tt, tol)
