"""GPU: round-2 kernels and semantics -- the nine-lane gather against the lane-per-slot one, the
single-reduction PCG against the classic recurrences, solver exit codes, quadrature tables of
contexts that live side by side, and the residual's rounding floor at small strain."""
import numpy as np
import pytest

import fea_gpu as fg
from conftest import block_model, load_brick_fine, load_golden
from oracle.oracle import PortOracle
from test_gpu_parity import RTOL_ELEM, RTOL_SOLVE, deformed, make_gpu, relmax

pytestmark = pytest.mark.gpu


GATHER_VARIANTS = {
    # name: (gather_mode, gather_sym, chunk_tiles, chunk_overlap)
    "every slot its own list": (1, 0, 0, 1),
    "upper triangle + mirror (default)": (1, 1, 0, 1),
    "nine lanes per block": (9, 0, 0, 1),
    "upper + mirror, chunks of 3 tiles": (1, 1, 3, 0),
    "upper + mirror, chunks of 7 tiles, gather beside the next chunk": (1, 1, 7, 1),
    "direct (push) assembly": (2, 1, 0, 1),
    "direct assembly, chunks of 5 tiles": (2, 1, 5, 1),
}


@pytest.mark.parametrize("case", ["neohook_brick", "kuhn6", "brick_fine"])
def test_gather_modes_agree_bitwise(case):
    """Every assembly variant sums a slot's contributions in the same order (ascending element id,
    fea_solver.c:873-883): identical bits in K, with and without the Dirichlet cancellation folded in
    (:1244-1257), whether a lane gathers every slot itself, only the upper triangle (the block goes to the mirror
    slot transposed), the element kernel writes destination-ordered cells, or the work is cut into chunks."""
    if case == "kuhn6":
        m = block_model((6, 7, 5), model=0, bc_style=1, dy=0.01)
        x = deformed(m, 4, 0.004)
    elif case == "brick_fine":
        m, _, x = load_brick_fine()
    else:
        m, _ = load_golden(case)
        x = deformed(m, 5)
    g = make_gpu(m)
    g.set_nodes(x)
    vals = {}
    for name, (mode, sym, chunk, overlap) in GATHER_VARIANTS.items():
        for k, v in (("gather_mode", mode), ("gather_sym", sym), ("chunk_tiles", chunk), ("chunk_overlap", overlap)):
            g.set_param(k, v)
        g.assemble_all(True)
        plain = g.get_csr()[3].copy()
        ke = g.element_matrix(len(m.conn) // 2)
        g.assemble_all(True, fuse_bc=True)
        fused = g.get_csr()[3].copy()
        g.assemble_stiffness()
        vals[name] = (plain, fused, g.get_csr()[3].copy(), ke)
    g.close()
    ref = vals["every slot its own list"]
    for name, v in vals.items():
        for a, b in zip(ref, v):
            assert np.array_equal(a, b), name
    assert np.array_equal(ref[0], ref[2])
    assert not np.array_equal(ref[0], ref[1])        # the cancellation did something


@pytest.mark.parametrize("x0", [fg.X0_ZERO, fg.X0_RHS])
def test_single_reduction_pcg_matches_classic(x0):
    """Chronopoulos-Gear recurrences (one reduction per iteration, the multi-rank default) give the
    classic iterates up to rounding: same solution, iteration counts within a few."""
    m = block_model((6, 12, 6), model=1, bc_style=1, dy=0.02, box=(1.0, 2.0, 1.0))
    g, o = make_gpu(m), PortOracle(m)
    for s in (g, o):
        s.apply_increment(1.0); s.update_state(); s.assemble_stiffness(); s.assemble_residual(); s.apply_bc(0.0)
    o.solve_slae()
    g.set_param("pcg_variant", 0)
    it0, rr0, ok0 = g.solve(1e-13, 20000, x0)
    u0 = g.get_solution()
    g.set_param("pcg_variant", 1)
    it1, rr1, ok1 = g.solve(1e-13, 20000, x0)
    u1 = g.get_solution()
    p = g.phase_ms()
    assert ok0 and ok1 and rr0 <= 1e-13 and rr1 <= 1e-13 and p["pcg_exit"] == 1
    assert abs(it1 - it0) <= max(3, it0 // 50), (it0, it1)
    assert relmax(u1, u0) < RTOL_SOLVE and relmax(u1, o.get_solution()) < RTOL_SOLVE
    it2, rr2, ok2 = g.solve(1e-13, 20000, x0)                  # deterministic: same count, same bits
    assert it2 == it1 and np.array_equal(g.get_solution(), u1)
    it3, rr3, ok3 = g.solve(1e-13, 7, x0, allow_unconverged=True)
    assert not ok3 and it3 == 7                                # max_iter honoured and reported


@pytest.mark.parametrize("variant", [0, 1])
def test_stalled_solve_returns_its_own_code(variant):
    """A solve that ends on the stall / divergence guard is not a success (fea_solver.c:260-298 treats a
    failed solve as an error): FEA_GPU_ERR_STALLED unless the caller accepts the checkpoint."""
    m, _ = load_golden("neohook_brick")
    g = make_gpu(m)
    g.apply_increment(1.0); g.assemble_all(True); g.apply_bc(0.0)
    g.set_param("pcg_variant", variant)
    g.set_param("pcg_stall", 3)                                # three iterations without a new best ||r|| = "stalled"
    with pytest.raises(fg.FeaGpuError) as e:
        g.solve(1e-14, 20000)
    assert e.value.code == fg.ERR_STALLED
    it, rr, ok = g.solve(1e-14, 20000, accept_stall=True)
    assert ok and rr > 1e-14 and g.phase_ms()["pcg_exit"] == 2
    u_ck = g.get_solution()
    R = g.get_forces()
    assert np.linalg.norm(g.spmv(u_ck) - R) <= 1.5 * rr * np.linalg.norm(R)   # u is the iterate `relres` describes
    g.set_param("pcg_stall", 0)
    it, rr, ok = g.solve(1e-14, 20000)
    assert ok and rr <= 1e-14 and g.phase_ms()["pcg_exit"] == 1


def test_four_and_five_point_contexts_side_by_side():
    """Two live contexts on one device with different quadrature rules must not share tables
    (gauss_nodes4/5_tetr10, fea_solver.c:32-54)."""
    m5, _ = load_golden("neohook_brick")
    m4, _ = load_golden("neohook_brick")
    m4.gauss = 4
    x = deformed(m5, 6)
    g5, g4 = make_gpu(m5), make_gpu(m4)                         # the 4-point context is created last
    o5, o4 = PortOracle(m5), PortOracle(m4)
    for s in (g5, g4, o5, o4):
        s.set_nodes(x)
    for s in (g5, g4, g5, g4):                                  # interleaved passes
        s.update_state(); s.assemble_stiffness(); s.assemble_residual()
    for s in (o5, o4):
        s.update_state(); s.assemble_stiffness(); s.assemble_residual()
    assert relmax(g5.get_csr()[3], o5.get_csr()[2]) < RTOL_ELEM
    assert relmax(g4.get_csr()[3], o4.get_csr()[2]) < RTOL_ELEM
    assert relmax(g5.get_forces(), o5.get_forces()) < RTOL_ELEM
    assert relmax(g4.get_forces(), o4.get_forces()) < RTOL_ELEM


@pytest.mark.parametrize("model", [0, 1])
def test_small_strain_residual_floor(model):
    """At 1e-6 strain |sigma| ~ 1e-4 mu: the residual must carry rounding of size eps |sigma|, as the
    reference's direct sigma . grad N sum does (fea_solver.c:1094-1109), not eps mu."""
    m = block_model((3, 3, 3), model=model, bc_style=1, dy=0.0)
    rng = np.random.default_rng(2)
    x = m.nodes + 1e-6 * rng.standard_normal(m.nodes.shape)
    g, o = make_gpu(m), PortOracle(m)
    for s in (g, o):
        s.set_nodes(x); s.update_state(); s.assemble_residual()
    Rg, Ro = g.get_forces(), o.get_forces()
    assert np.abs(Ro).max() < 1e-2                              # small-strain state indeed
    # F itself (through F^-1 with absolute coordinates, :1141-1152) carries eps |X| / h ~ 1e-15, i.e. ~1e-10 of
    # a 6e-6 displacement gradient, on both sides: that, not the residual formula, sets the floor here
    assert relmax(Rg, Ro) < 1e-8
    assert relmax(g.get_state()[1], o.get_state()[1]) < 1e-8


def test_pcg_ilu_setting_is_chebyshev_jacobi():
    """The PCG_ILU request of a task file (fea_solver.c:260-280) selects a genuinely stronger
    preconditioner -- a fixed Chebyshev polynomial in D^-1 A: at least 3x fewer iterations, the same u."""
    m = block_model((8, 16, 8), model=1, bc_style=2, dy=0.01, box=(1.0, 2.0, 1.0))
    g, o = make_gpu(m), PortOracle(m)
    for s in (g, o):
        s.apply_increment(1.0); s.update_state(); s.assemble_stiffness(); s.assemble_residual(); s.apply_bc(0.0)
    o.solve_slae()
    it0, rr0, ok0 = g.solve(1e-13, 50000)
    u0 = g.get_solution()
    g.set_param("precond", 1)
    it1, rr1, ok1 = g.solve(1e-13, 50000)
    u1 = g.get_solution()
    it2, _, _ = g.solve(1e-13, 50000)
    assert ok0 and ok1 and rr1 <= 1e-13 and g.phase_ms()["pcg_exit"] == 1
    assert it0 >= 3 * it1, (it0, it1)
    assert relmax(u1, u0) < RTOL_SOLVE and relmax(u1, o.get_solution()) < RTOL_SOLVE
    assert it2 == it1 and np.array_equal(g.get_solution(), u1)          # deterministic
    g.set_param("precond", 0)
    it3, _, _ = g.solve(1e-13, 50000)
    assert it3 == it0
