"""Host C layer (fea-large_b200/host): the reference's public API over the C-ABI.
CPU part: self test, .sexp reader, shape functions.  GPU part: the feasolver_b200 binary
end to end against the reference's own Gmsh export (golden)."""
import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, block_model, load_golden, write_sexp
from oracle.oracle import PortOracle, load_sexp

HOST_LIB = os.path.join(ROOT, "fea-large_b200", "lib", "libfea_host.so")
BIN = os.path.join(ROOT, "fea-large_b200", "bin", "feasolver_b200")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_LIB) or not os.path.exists(BIN):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "fea-large_b200"), "all"], check=True)
    return C.CDLL(HOST_LIB)


class _Nodes(C.Structure):
    _fields_ = [("count", C.c_int), ("rows", C.POINTER(C.POINTER(C.c_double)))]


class _Elems(C.Structure):
    _fields_ = [("count", C.c_int), ("rows", C.POINTER(C.POINTER(C.c_int)))]


class _Presc(C.Structure):
    _fields_ = [("node", C.c_int), ("values", C.c_double * 3), ("type", C.c_int)]


class _PrescArr(C.Structure):
    _fields_ = [("count", C.c_int), ("items", C.POINTER(_Presc))]


class _Model(C.Structure):
    _fields_ = [("model", C.c_int), ("parameters", C.c_double * 10), ("parameters_count", C.c_int)]


class _Task(C.Structure):
    _fields_ = [("type", C.c_int), ("model", _Model), ("solver_type", C.c_int), ("solver_tolerance", C.c_double),
                ("solver_max_iter", C.c_int), ("dof", C.c_int), ("ele_type", C.c_int),
                ("load_increments_count", C.c_int), ("desired_tolerance", C.c_double),
                ("max_newton_count", C.c_int), ("linesearch_max", C.c_int), ("arclength_max", C.c_int),
                ("modified_newton", C.c_int), ("export_file", C.c_char_p)]


class _Params(C.Structure):
    _fields_ = [("nodes_per_element", C.c_int), ("gauss_nodes_count", C.c_int)]


def c_load(host, path):
    t, p, n, e, b = (C.c_void_p() for _ in range(5))
    ok = host.sexp_data_load(path.encode(), C.byref(t), C.byref(p), C.byref(n), C.byref(e), C.byref(b))
    if not ok:
        return None
    task = C.cast(t, C.POINTER(_Task)).contents
    par = C.cast(p, C.POINTER(_Params)).contents
    na = C.cast(n, C.POINTER(_Nodes)).contents
    ea = C.cast(e, C.POINTER(_Elems)).contents
    pa = C.cast(b, C.POINTER(_PrescArr)).contents
    nodes = np.ctypeslib.as_array(na.rows[0], shape=(na.count, 3)).copy()      # rows[0] = the flat block
    conn = np.ctypeslib.as_array(ea.rows[0], shape=(ea.count, par.nodes_per_element)).copy()
    presc = [(pa.items[i].node, pa.items[i].type, tuple(pa.items[i].values)) for i in range(pa.count)]
    return task, par, nodes, conn, presc


def test_startup_self_test(host):
    assert host.do_tests() == 1          # solver-large/tests.c known answers


def test_sexp_reader_round_trip(host, tmp_path):
    m = block_model((2, 3, 2), model=1, dy=0.05)
    m.desired_tolerance, m.max_newton, m.solver_type, m.modified_newton = 1e-6, 110, 2, True
    for shuffle in (False, True):
        path = str(tmp_path / f"block{int(shuffle)}.sexp")
        write_sexp(path, m, load_increments=7, shuffle_keys=shuffle)
        task, par, nodes, conn, presc = c_load(host, path)
        assert np.array_equal(nodes, m.nodes) and np.array_equal(conn, m.conn)
        assert [p[0] for p in presc] == list(m.presc_node) and [p[1] for p in presc] == list(m.presc_type)
        assert np.array_equal(np.array([p[2] for p in presc]), m.presc_vals)
        assert task.model.model == 1 and task.model.parameters[0] == 100.0 and task.model.parameters[1] == 100.0
        assert task.load_increments_count == 7 and task.max_newton_count == 110 and task.modified_newton == 1
        assert task.desired_tolerance == 1e-6 and task.solver_type == 2 and task.solver_tolerance == 1e-14
        assert par.gauss_nodes_count == 5 and par.nodes_per_element == 10
        py = load_sexp(path)              # the test-side Python reader agrees
        assert np.array_equal(py.nodes, nodes) and np.array_equal(py.conn, conn)


def test_sexp_reader_rejects_broken_input(host, tmp_path):
    m = block_model((1, 1, 1))
    good = str(tmp_path / "good.sexp")
    write_sexp(good, m)
    text = open(good).read()
    cases = {"nolambda": text.replace(":lambda 100", ""), "shortrow": re.sub(r"\((\d+ ){9}\d+\)", "(0 1 2)", text, count=1),
             "notask": text.replace("(task", "(job"), "nosolutionkey": text.replace(":load-increments-count 2", "")}
    for name, body in cases.items():
        path = str(tmp_path / f"{name}.sexp")
        open(path, "w").write(body)
        assert c_load(host, path) is None, name
    assert c_load(host, str(tmp_path / "missing.sexp")) is None


@pytest.mark.skipif(not os.path.isdir("/root/reference/solver-large/data"), reason="reference data not present")
@pytest.mark.parametrize("name", ["neohook_brick", "a5_brick_analytical", "brick_fine"])
def test_sexp_reader_on_shipped_models(host, name):
    path = f"/root/reference/solver-large/data/{name}.sexp"
    task, par, nodes, conn, presc = c_load(host, path)
    py = load_sexp(path)
    assert np.array_equal(nodes, py.nodes) and np.array_equal(conn, py.conn)
    assert [p[0] for p in presc] == list(py.presc_node)
    assert task.model.model == py.model and task.solver_type == py.solver_type
    assert task.load_increments_count == py.load_increments and task.max_newton_count == py.max_newton


def test_host_shape_functions_match_oracle_tables(host):
    host.tetrahedra10_isoform.restype = C.c_double
    host.tetrahedra10_isoform.argtypes = [C.c_int] + [C.c_double] * 3
    host.tetrahedra10_disoform.restype = C.c_double
    host.tetrahedra10_disoform.argtypes = [C.c_int, C.c_int] + [C.c_double] * 3
    gt, N, dN = PortOracle.tables(5)
    for g in range(5):
        r, s, t = gt[g, 1:]
        assert np.allclose([host.tetrahedra10_isoform(a, r, s, t) for a in range(10)], N[g], atol=1e-15)
        for d in range(3):
            assert np.allclose([host.tetrahedra10_disoform(a, d, r, s, t) for a in range(10)], dN[g, d], atol=1e-14)


def _numbers(line):
    return [float(x) for x in line.split()]


@pytest.mark.gpu
def test_feasolver_binary_reproduces_reference_export(host, tmp_path):
    """bin/feasolver_b200 on the shipped analytical brick, two increments: same log structure
    and the same Gmsh file as the reference binary's own export (golden), to the 1e-6 the
    %f format carries."""
    m, z = load_golden("neohook_brick_analytical")
    path = str(tmp_path / "neohook_brick_analytical.sexp")
    write_sexp(path, m, load_increments=2)
    run = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    out = run.stdout
    assert "test_matrix result: *pass*" in out
    assert out.count("Newton iteration") == int(z["newton_count"])       # 25 linear solves in the reference
    tol = [float(x) for x in re.findall(r"Tolerance <X,R> = (\S+)", out)]
    assert np.allclose(tol, z["newton_tol"], rtol=1e-5, atol=1e-12)
    assert out.count("Load increment") == 2
    got = open(str(tmp_path / "neohook_brick_analytical.msh")).read().splitlines()
    want = gzip.open(os.path.join(GOLDEN, "neohook_brick_analytical_2steps.msh.gz"), "rt").read().splitlines()
    assert len(got) == len(want)
    worst = 0.0
    for a, b in zip(got, want):
        if a == b:
            continue
        na, nb = _numbers(a), _numbers(b)
        assert len(na) == len(nb), (a, b)
        worst = max(worst, max(abs(x - y) for x, y in zip(na, nb)))
    assert worst <= 2e-6, worst


@pytest.mark.gpu
def test_feasolver_reports_nonconvergence_like_the_reference(host, tmp_path):
    """max-newton-count reached: step rolled back, error logged, export still written
    (fea_solver.c:225-239)."""
    m, _ = load_golden("neohook_brick")
    m.max_newton = 2
    path = str(tmp_path / "stuck.sexp")
    write_sexp(path, m, load_increments=3)
    run = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert run.returncode == 0
    assert "Unable to finish load step in 2 Newton iterations,exit" in run.stdout
    assert "Load increment 1 finished" in run.stdout and "Load increment 2 finished" not in run.stdout
    msh = open(str(tmp_path / "stuck.msh")).read()
    # current_load_step is rolled back to -1, so the reference's export loop (load <= step,
    # :1440) writes the mesh and no data sections at all
    assert msh.count("$Nodes") == 1 and msh.count("$NodeData") == 0


def _msh_sections(path):
    """{(kind, step): array} of the $NodeData / $ElementData blocks of a Gmsh 2.0 file."""
    out, lines, i = {}, open(path).read().splitlines(), 0
    while i < len(lines):
        if lines[i] in ("$NodeData", "$ElementData"):
            kind, step, n = lines[i], int(lines[i + 6]), int(lines[i + 8])
            out[(kind, step)] = np.array([_numbers(l)[1:] for l in lines[i + 9:i + 9 + n]])
            i += 9 + n
        else:
            i += 1
    return out


@pytest.mark.gpu
def test_feasolver_halves_an_increment_that_inverts_elements(host, tmp_path):
    """One increment of three element heights on a uniaxial bar: the boundary move alone folds the top layer
    (mid-side nodes end up beyond the quarter point), where the reference's log(det F) turns NaN
    (fea_model.c:105).  solve() rolls the increment back and applies it in halves; the converged state is
    the homogeneous one of exact-solutions/uniaxial at the full stretch."""
    from oracle.oracle import uniaxial_neohookean
    m = block_model((2, 2, 2), model=1, bc_style=2, dy=1.5)
    m.desired_tolerance, m.max_newton, m.solver_type, m.modified_newton = 1e-14, 60, 0, False
    path = str(tmp_path / "fold.sexp")
    write_sexp(path, m, load_increments=1)
    run = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "Gauss points inverted by a load fraction of 1, halving it" in run.stdout
    assert "Load increment 1 finished" in run.stdout and "Unable to finish" not in run.stdout
    sec = _msh_sections(str(tmp_path / "fold.msh"))
    u, sig = sec[("$NodeData", 1)], sec[("$ElementData", 1)]
    k2, syy = uniaxial_neohookean(2.5)
    assert abs(u[:, 1].max() - 1.5) < 2e-6                       # the whole increment was applied
    assert np.abs(sig[:, 4] - syy).max() < 1e-4 * syy            # %f carries six decimals
    assert np.abs(sig[:, [0, 8]]).max() < 1e-4                   # lateral faces stress free


@pytest.mark.gpu
def test_feasolver_line_search_as_the_prototype(host, tmp_path):
    """:line-search :max 3 -- golden-section probes of |eta <u, R(x + eta u)>| on [0.5, 1]
    (solver-prototype/cartesian3d/large/cartesian3d_large.m:85-119); with :max 0 solve() is the shipped one.
    Both runs must land on the same equilibrium."""
    outs = {}
    for ls in (0, 3):
        m, _ = load_golden("neohook_brick")
        m.desired_tolerance, m.max_newton, m.modified_newton = 1e-12, 60, False
        m.extra["linesearch"] = ls
        path = str(tmp_path / f"ls{ls}.sexp")
        write_sexp(path, m, load_increments=1)
        run = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
        assert run.returncode == 0, run.stderr[-2000:]
        assert ("Line search: eta" in run.stdout) == (ls > 0)
        assert "Unable to finish" not in run.stdout
        outs[ls] = _msh_sections(str(tmp_path / f"ls{ls}.msh"))[("$NodeData", 1)]
    assert np.abs(outs[0] - outs[3]).max() <= 2e-6


@pytest.mark.gpu
def test_feasolver_without_per_increment_snapshots(host, tmp_path):
    """FEA_KEEP_STEPS=0: only the last increment is pulled to the host and exported (fea_solver.c:605-636 keeps all)."""
    m, _ = load_golden("a5_brick")
    path = str(tmp_path / "keep.sexp")
    write_sexp(path, m, load_increments=2)
    full = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    a = _msh_sections(str(tmp_path / "keep.msh"))
    last = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600,
                          env=dict(os.environ, FEA_KEEP_STEPS="0"))
    b = _msh_sections(str(tmp_path / "keep.msh"))
    assert full.returncode == 0 and last.returncode == 0
    assert set(a) == {(k, s) for k in ("$NodeData", "$ElementData") for s in (0, 1, 2)}
    assert set(b) == {(k, s) for k in ("$NodeData", "$ElementData") for s in (0, 2)}
    assert np.array_equal(a[("$NodeData", 2)], b[("$NodeData", 2)])


@pytest.mark.gpu
def test_feasolver_secant_predictor_reaches_the_same_equilibria(host, tmp_path):
    """FEA_PREDICTOR=1: increments after the first start from x_k + (x_k - x_{k-1}) (fea_gpu_extrapolate_nodes) instead
    of the boundary move alone (fea_solver.c:168).  Same exported states, fewer Newton iterations."""
    m = block_model((3, 4, 3), model=1, bc_style=2, dy=0.1)
    m.desired_tolerance, m.max_newton, m.solver_type, m.modified_newton = 1e-12, 60, 0, False
    path = str(tmp_path / "pred.sexp")
    write_sexp(path, m, load_increments=5)
    out, newton = {}, {}
    for p in ("0", "1"):
        run = subprocess.run([BIN, path], capture_output=True, text=True, cwd=str(tmp_path), timeout=600,
                             env=dict(os.environ, FEA_PREDICTOR=p))
        assert run.returncode == 0, run.stderr[-2000:]
        assert "Unable to finish" not in run.stdout and "inverted" not in run.stdout
        out[p] = _msh_sections(str(tmp_path / "pred.msh"))
        newton[p] = run.stdout.count("Newton iteration")
    for key in out["0"]:
        assert np.abs(out["0"][key] - out["1"][key]).max() <= 3e-6, key      # %f carries six decimals
    assert newton["1"] < newton["0"], newton
