"""CPU: pin the plain-C oracle (oracle/oracle_fea.c) against the committed golden vectors,
which were produced by the reference's own compiled code (tests/golden/make_golden.py),
against the reference's known-answer test (solver-large/tests.c:17-22) and against the
closed forms of exact-solutions/uniaxial."""
import numpy as np
import pytest

from conftest import GOLDEN, csr_mv
from oracle.oracle import PortOracle, uniaxial_a5, uniaxial_neohookean


def test_tests_c_known_answers():
    # solver-large/tests.c:17-22: A, B and the three expected products
    A = np.array([[1, 2, 0], [2, 0, 3], [0, 2, 3]], float)
    B = np.array([[0, 2, 1], [1, 1, 1], [3, 2, -1]], float)
    assert np.array_equal(PortOracle.matmul(0, A, B), [[2, 4, 3], [9, 10, -1], [11, 8, -1]])
    assert np.array_equal(PortOracle.matmul(1, A, B), [[2, 4, 3], [6, 8, 0], [12, 9, 0]])
    assert np.array_equal(PortOracle.matmul(2, A, B), [[4, 3, 7], [3, 5, 3], [7, 5, 1]])


def test_tables_and_models_match_reference_compiled():
    z = np.load(f"{GOLDEN}/model_tables.npz")
    for ng in (4, 5):
        gt, N, dN = PortOracle.tables(ng)
        assert np.array_equal(gt, z[f"gauss{ng}"])
        assert np.array_equal(N, z[f"N{ng}"])
        assert np.array_equal(dN, z[f"dN{ng}"])
        assert np.allclose(N.sum(axis=1), 1.0, atol=1e-15)        # partition of unity (test_isoform.m)
        assert np.allclose(dN.sum(axis=2), 0.0, atol=1e-14)
    for model, tag in ((0, "a5"), (1, "nh")):
        for F, S, Ct in zip(z["F"], z["S_" + tag], z["C_" + tag]):
            s, c = PortOracle.model_eval(model, 100.0, 100.0, F)
            assert np.array_equal(s, S) and np.array_equal(c, Ct)


def test_five_point_rule_has_negative_centre_weight():
    gt, _, _ = PortOracle.tables(5)
    assert gt[0, 0] == (-4 / 5.) / 6. and np.isclose(gt[:, 0].sum(), 1 / 6.)
    gt4, _, _ = PortOracle.tables(4)
    assert gt4[0, 1] == 0.58541020 and gt4[0, 2] == 0.13819660   # the reference's 8-digit literals


def test_element_phase_bitwise(brick):
    name, m, z = brick
    o = PortOracle(m)
    o.apply_increment(1.0)
    o.update_state()
    F, S = o.get_state()
    g, detJ = o.get_gradients()
    assert np.array_equal(F, z["F"]) and np.array_equal(S, z["S"]) and np.array_equal(detJ, z["detJ"])
    assert np.array_equal(g[z["probe_elems"]], z["g_sample"])
    for k, e in enumerate(z["probe_elems"]):
        assert np.array_equal(o.element_matrix(int(e)), z["Ke"][k])


def test_global_assembly_bc_and_solve(brick):
    name, m, z = brick
    o = PortOracle(m)
    o.apply_increment(1.0)
    o.update_state()
    o.assemble_stiffness()
    o.assemble_residual()
    rp, ci, v = o.get_csr()
    assert len(v) == int(z["nnz"]) == 145737                        # C1 nonzero count
    assert np.array_equal(rp, z["rowptr"]) and np.array_equal(ci[:2000], z["colidx_head"])
    assert np.all(np.diff(ci)[np.diff(np.repeat(np.arange(len(rp) - 1), np.diff(rp))) == 0] > 0)  # sorted rows
    assert np.array_equal(o.get_forces(), z["R"])
    for p, kv in zip(z["probes"], z["Kv"]):
        assert np.allclose(csr_mv(rp, ci, v, p), kv, rtol=0, atol=1e-13 * np.abs(kv).max())
    diag = np.array([v[rp[i]:rp[i + 1]][ci[rp[i]:rp[i + 1]] == i][0] for i in range(m.n_dof)])
    assert np.array_equal(diag, z["Kdiag"])
    o.apply_bc(0.0)
    rp, ci, v = o.get_csr()
    assert np.array_equal(o.get_forces(), z["R_bc"])
    for p, kv in zip(z["probes"], z["Kv_bc"]):
        assert np.allclose(csr_mv(rp, ci, v, p), kv, rtol=0, atol=1e-13 * np.abs(kv).max())
    o.solve_slae()
    assert np.array_equal(o.get_solution(), z["u_first"])


def test_newton_driver_matches_reference_solve(brick):
    name, m, z = brick
    o = PortOracle(m)
    done, tu, tt = o.newton_solve(2, m.desired_tolerance, m.modified_newton, m.max_newton)
    assert done == 2 and len(tu) == int(z["newton_count"])
    assert np.array_equal(tt, z["newton_tol"])
    assert np.array_equal(tu[:3], z["newton_u_head"])
    assert np.allclose(tu.sum(axis=0), z["newton_u_sum"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name,closed", [("neohook_brick_analytical", uniaxial_neohookean),
                                         ("a5_brick_analytical", uniaxial_a5)])
def test_uniaxial_closed_form(name, closed):
    """exact-solutions/uniaxial: quadratic tets reproduce the homogeneous state exactly once
    Newton is driven to a tight tolerance."""
    from conftest import load_golden
    m, _ = load_golden(name)
    o = PortOracle(m)
    o.newton_solve(1, 1e-22, False, 60)
    F, S = o.get_state()
    k1 = 1 + 0.05 / 6
    k2, sig = closed(k1)
    assert np.allclose(S[:, :, 1, 1], sig, rtol=1e-9)
    assert np.allclose(F[:, :, 1, 1], k1, rtol=1e-10) and np.allclose(F[:, :, 0, 0], k2, rtol=1e-9)


def test_closed_form_table_values():
    # BASELINE.md section 4
    for s, k2e, se in [(1, 0.997925301, 2.079482974), (2, 0.995867686, 4.151485076), (3, 0.993826886, 6.216327914)]:
        k2, sig = uniaxial_neohookean(1 + s * 0.05 / 6)
        assert abs(k2 - k2e) < 1e-9 and abs(sig - se) < 1e-8
    for s, k2e, se in [(1, 0.997905793, 2.118310408), (2, 0.995789748, 4.307607909), (3, 0.993651725, 6.569473018)]:
        k2, sig = uniaxial_a5(1 + s * 0.05 / 6)
        assert abs(k2 - k2e) < 1e-9 and abs(sig - se) < 1e-8


def test_brick_fine_unstructured_mesh_matches_reference_compiled():
    """C1f: 102 210 DOF, 8 300 196 nonzeros (81.2 per row, max 330) -- SURVEY section 8."""
    from conftest import load_brick_fine
    m, z, x = load_brick_fine()
    o = PortOracle(m)
    o.set_nodes(x); o.update_state(); o.assemble_stiffness(); o.assemble_residual()
    rp, ci, v = o.get_csr()
    assert len(v) == int(z["nnz"]) == 8300196 and int(np.diff(rp).max()) == int(z["max_row"]) == 330
    probes = np.random.default_rng(int(z["probe_seed"])).standard_normal((2, m.n_dof))
    for p, ks, kn in zip(probes, z["Kv_sample"], z["Kv_norm"]):
        kv = csr_mv(rp, ci, v, p)
        assert np.allclose(kv[::40], ks, rtol=0, atol=1e-12 * np.abs(ks).max()) and abs(np.linalg.norm(kv) - kn) <= 1e-12 * kn
    R = o.get_forces()
    assert np.array_equal(R[::40], z["R_sample"]) and abs(np.linalg.norm(R) - float(z["R_norm"])) <= 1e-13 * float(z["R_norm"])
    assert np.array_equal(o.get_state()[1][::997], z["S_sample"])
