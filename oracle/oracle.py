"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the two CPU checkers:

* ``RefOracle``  -- the reference's own C objects compiled from /root/reference
  into ``oracle/_ref/libfea_ref.so`` (see oracle/Makefile, ref_standin.c).
* ``PortOracle`` -- the plain-C restatement ``oracle/oracle_fea.c``.

Both expose the same methods so tests can run either against the CUDA path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module; nothing under fea-large_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libfea_ref.so")
PORT_LIB = os.path.join(HERE, "_build", "libfea_oracle.so")

MODEL_A5, MODEL_NH = 0, 1
SOLVER_CG, SOLVER_PCG_ILU, SOLVER_CHOLESKY = 0, 1, 2

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True, port: bool = True) -> None:
    """Compile the checkers (gcc only).  `ref` is a no-op without /root/reference."""
    targets = [t for t, on in (("port", port), ("ref", ref)) if on]
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


# ---------------------------------------------------------------------------
# model files (.sexp) -- grammar as used by the reference loader
# (solver-large/sexp_loader.c:32-247); test-side reader only


@dataclass
class Model:
    nodes: np.ndarray            # [n][3] float64
    conn: np.ndarray             # [e][10] int32
    presc_node: np.ndarray       # [p] int32
    presc_type: np.ndarray       # [p] int32 (bitmask 1=x 2=y 4=z)
    presc_vals: np.ndarray       # [p][3] float64
    model: int = MODEL_A5
    lam: float = 100.0
    mu: float = 100.0
    gauss: int = 5
    load_increments: int = 0
    desired_tolerance: float = 1e-8
    modified_newton: bool = True
    max_newton: int = 0
    solver_type: int = SOLVER_CG
    solver_tolerance: float = 1e-14
    solver_max_iter: int = 20000
    extra: dict = field(default_factory=dict)

    @property
    def n_dof(self) -> int:
        return 3 * len(self.nodes)


_TOKEN = re.compile(r"""\(|\)|"[^"]*"|[^\s()]+""")


def _parse_sexp(text: str):
    text = re.sub(r";[^\n]*", "", text)
    stack, cur = [], []
    for tok in _TOKEN.findall(text):
        if tok == "(":
            stack.append(cur)
            cur = []
        elif tok == ")":
            done = cur
            cur = stack.pop()
            cur.append(done)
        else:
            cur.append(tok)
    return cur[0]


def _attr(form, key, default=None):
    for i, t in enumerate(form):
        if isinstance(t, str) and t.lower() == ":" + key and i + 1 < len(form):
            return form[i + 1]
    return default


def _walk(form):
    yield form
    for t in form:
        if isinstance(t, list):
            yield from _walk(t)


def load_sexp(path: str) -> Model:
    with open(path) as f:
        tree = _parse_sexp(f.read())
    assert tree[0].lower() == "task"
    kw = {}
    nodes = conn = None
    pn, pt, pv = [], [], []
    for form in _walk(tree):
        if not form or not isinstance(form[0], str):
            continue
        head = form[0].lower()
        if head == "model":
            name = _attr(form, "name", "A5").upper()
            kw["model"] = MODEL_NH if name == "COMPRESSIBLE_NEOHOOKEAN" else MODEL_A5
        elif head == "model-parameters":
            kw["lam"] = float(_attr(form, "lambda"))
            kw["mu"] = float(_attr(form, "mu"))
        elif head == "solution":
            kw["desired_tolerance"] = float(_attr(form, "desired-tolerance"))
            kw["load_increments"] = int(_attr(form, "load-increments-count"))
            kw["modified_newton"] = _attr(form, "modified-newton", "no").upper() in ("YES", "TRUE")
            kw["max_newton"] = int(_attr(form, "max-newton-count"))
        elif head == "slae-solver":
            kw["solver_type"] = {"CG": 0, "PCG_ILU": 1, "CHOLESKY": 2}[_attr(form, "type", "CG").upper()]
            if _attr(form, "tolerance") is not None:
                kw["solver_tolerance"] = float(_attr(form, "tolerance"))
            if _attr(form, "max-iterations") is not None:
                kw["solver_max_iter"] = int(_attr(form, "max-iterations"))
        elif head == "element-type":
            kw["gauss"] = int(_attr(form, "gauss-nodes-count"))
        elif head == "nodes":
            nodes = np.array([[float(v) for v in row] for row in form[1:]], dtype=np.float64)
        elif head == "elements":
            conn = np.array([[int(v) for v in row] for row in form[1:]], dtype=np.int32)
        elif head == "presc-node":
            pn.append(int(_attr(form, "node-id")))
            pt.append(int(_attr(form, "type")))
            pv.append([float(_attr(form, "x")), float(_attr(form, "y")), float(_attr(form, "z"))])
    return Model(nodes=np.ascontiguousarray(nodes), conn=np.ascontiguousarray(conn),
                 presc_node=np.array(pn, dtype=np.int32), presc_type=np.array(pt, dtype=np.int32),
                 presc_vals=np.array(pv, dtype=np.float64).reshape(-1, 3), **kw)


# ---------------------------------------------------------------------------


class _Base:
    """Shared method surface; subclasses bind the symbol prefix."""

    _prefix = ""
    _lib = None

    def _f(self, name):
        return getattr(self._lib, self._prefix + name)

    def __init__(self, m: Model):
        self.m = m
        self.ne, self.ng, self.n = len(m.conn), m.gauss, m.n_dof
        create = self._f("create")
        create.restype = C.c_void_p
        create.argtypes = [C.c_int, _dp, C.c_int, _ip, C.c_int, _ip, _ip, _dp,
                           C.c_int, C.c_double, C.c_double, C.c_int]
        pv = np.ascontiguousarray(m.presc_vals.reshape(-1, 3)) if len(m.presc_node) else np.zeros((1, 3))
        pn = m.presc_node if len(m.presc_node) else np.zeros(1, np.int32)
        pt = m.presc_type if len(m.presc_node) else np.zeros(1, np.int32)
        self.h = C.c_void_p(create(len(m.nodes), m.nodes, len(m.conn), m.conn, len(m.presc_node),
                                   pn, pt, pv, m.model, m.lam, m.mu, m.gauss))

    def close(self):
        if self.h:
            d = self._f("destroy")
            d.argtypes = [C.c_void_p]
            d.restype = None
            d(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args, argtypes=(), restype=None):
        f = self._f(name)
        f.argtypes = [C.c_void_p, *argtypes]
        f.restype = restype
        return f(self.h, *args)

    def set_nodes(self, x):
        self._call("set_nodes", np.ascontiguousarray(x, np.float64), argtypes=[_dp])

    def get_nodes(self):
        x = np.empty((self.n // 3, 3))
        self._call("get_nodes", x, argtypes=[_dp])
        return x

    def apply_increment(self, lam=1.0):
        self._call("apply_increment", float(lam), argtypes=[C.c_double])

    def update_state(self):
        self._call("update_state")

    def get_state(self):
        F = np.empty((self.ne, self.ng, 3, 3))
        S = np.empty((self.ne, self.ng, 3, 3))
        self._call("get_state", F, S, argtypes=[_dp, _dp])
        return F, S

    def get_gradients(self):
        g = np.empty((self.ne, self.ng, 3, 10))
        d = np.empty((self.ne, self.ng))
        self._call("get_gradients", g, d, argtypes=[_dp, _dp])
        return g, d

    def element_matrix(self, e, part=0):
        ke = np.empty((30, 30))
        self._call("element_matrix", int(e), ke, int(part), argtypes=[C.c_int, _dp, C.c_int])
        return ke

    def apply_bc(self, lam=0.0):
        self._call("apply_bc", float(lam), argtypes=[C.c_double])

    def get_csr(self):
        nnz = self._call("nnz", restype=C.c_long)
        rp = np.empty(self.n + 1, np.int32)
        ci = np.empty(nnz, np.int32)
        v = np.empty(nnz)
        self._call("get_csr", rp, ci, v, argtypes=[_ip, _ip, _dp])
        return rp, ci, v

    def get_forces(self):
        r = np.empty(self.n)
        self._call("get_forces", r, argtypes=[_dp])
        return r

    def set_forces(self, r):
        self._call("set_forces", np.ascontiguousarray(r, np.float64), argtypes=[_dp])

    def get_solution(self):
        u = np.empty(self.n)
        self._call("get_solution", u, argtypes=[_dp])
        return u

    def update_with_solution(self):
        self._call("update_with_solution")


class RefOracle(_Base):
    """The reference's own compiled element/BC/Newton code (kind = "reference")."""

    _prefix = "ref_"

    def __init__(self, m: Model):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libfea_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        type(self)._lib = C.CDLL(REF_LIB)
        super().__init__(m)

    def assemble_stiffness(self):
        self._call("stiffness")

    def assemble_residual(self):
        self._call("residual")

    def solve_slae(self):
        return self._call("solve_slae", restype=C.c_int)

    # -- statics ------------------------------------------------------------
    @staticmethod
    def lib():
        return C.CDLL(REF_LIB)

    @staticmethod
    def set_scatter_mode(mode: int):
        RefOracle.lib().ref_set_scatter_mode(int(mode))

    @staticmethod
    def model_eval(model, lam, mu, F):
        lib = RefOracle.lib()
        lib.ref_model_eval.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
        S, ct = np.empty((3, 3)), np.empty((3, 3, 3, 3))
        lib.ref_model_eval(model, lam, mu, np.ascontiguousarray(F, np.float64), S, ct)
        return S, ct

    @staticmethod
    def matmul(which, A, B):
        lib = RefOracle.lib()
        lib.ref_matmul.argtypes = [C.c_int, _dp, _dp, _dp]
        R = np.empty((3, 3))
        lib.ref_matmul(which, np.ascontiguousarray(A, np.float64), np.ascontiguousarray(B, np.float64), R)
        return R

    @staticmethod
    def tables(count):
        lib = RefOracle.lib()
        lib.ref_gauss_table.argtypes = [C.c_int, _dp]
        lib.ref_shape_tables.argtypes = [C.c_double] * 3 + [_dp, _dp]
        gt = np.empty((count, 4))
        lib.ref_gauss_table(count, gt)
        N, dN = np.empty((count, 10)), np.empty((count, 3, 10))
        for g in range(count):
            lib.ref_shape_tables(gt[g, 1], gt[g, 2], gt[g, 3], N[g], dN[g])
        return gt, N, dN

    @staticmethod
    def run_solve(m: Model, load_increments=None, desired_tol=None, modified_newton=None,
                  max_newton=None, msh_path="/tmp/fea_ref_out.msh"):
        """The reference's solve() end to end; returns (rhs[k][n], u[k][n], tol[k])."""
        lib = RefOracle.lib()
        lib.ref_run_solve.argtypes = [C.c_int, _dp, C.c_int, _ip, C.c_int, _ip, _ip, _dp, C.c_int,
                                      C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_int, C.c_int, C.c_double, C.c_int, C.c_char_p]
        li = m.load_increments if load_increments is None else load_increments
        k = lib.ref_run_solve(len(m.nodes), m.nodes, len(m.conn), m.conn, len(m.presc_node),
                              m.presc_node, m.presc_type, np.ascontiguousarray(m.presc_vals),
                              m.model, m.lam, m.mu, m.gauss, li,
                              m.desired_tolerance if desired_tol is None else desired_tol,
                              int(m.modified_newton if modified_newton is None else modified_newton),
                              m.max_newton if max_newton is None else max_newton,
                              m.solver_type, m.solver_tolerance, m.solver_max_iter,
                              msh_path.encode())
        lib.ref_trace_get.argtypes = [C.c_int, _dp, _dp]
        rhs, sol = np.empty((k, m.n_dof)), np.empty((k, m.n_dof))
        for i in range(k):
            lib.ref_trace_get(i, rhs[i], sol[i])
        nt = lib.ref_tolerance_count()
        tol = np.empty(max(nt, 1))
        lib.ref_tolerance_get.argtypes = [_dp]
        lib.ref_tolerance_get(tol)
        lib.ref_trace_reset()
        return rhs, sol, tol[:nt]


class PortOracle(_Base):
    """Plain-C restatement (kind = "port")."""

    _prefix = "orc_"

    def __init__(self, m: Model):
        if not os.path.exists(PORT_LIB):
            build(ref=False)
        type(self)._lib = C.CDLL(PORT_LIB)
        super().__init__(m)

    def assemble_stiffness(self):
        self._call("assemble_stiffness")

    def assemble_residual(self):
        self._call("assemble_residual")

    def solve_slae(self, rel_tol=1e-15, max_iter=200000):
        return self._call("solve_slae", rel_tol, max_iter, argtypes=[C.c_double, C.c_int], restype=C.c_int)

    def dot_forces_solution(self):
        return self._call("dot_forces_solution", restype=C.c_double)

    def newton_solve(self, load_increments, desired_tol, modified_newton, max_newton,
                     lin_tol=1e-15, lin_max_iter=200000, trace_cap=4096):
        tu = np.zeros((trace_cap, self.n))
        tt = np.zeros(trace_cap)
        nt = C.c_int(0)
        done = self._call("newton_solve", load_increments, desired_tol, int(modified_newton), max_newton,
                          lin_tol, lin_max_iter, tu, tt, trace_cap, C.byref(nt),
                          argtypes=[C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp,
                                    C.c_int, C.POINTER(C.c_int)], restype=C.c_int)
        k = min(nt.value, trace_cap)
        return done, tu[:k].copy(), tt[:k].copy()

    def time_assembly(self, reps=1):
        return self._call("time_assembly", reps, argtypes=[C.c_int], restype=C.c_double)

    @staticmethod
    def lib():
        if not os.path.exists(PORT_LIB):
            build(ref=False)
        return C.CDLL(PORT_LIB)

    @staticmethod
    def tables(count):
        lib = PortOracle.lib()
        lib.orc_gauss_table.argtypes = [C.c_int, _dp]
        lib.orc_shape_functions.argtypes = [C.c_double] * 3 + [_dp, _dp]
        gt = np.empty((count, 4))
        lib.orc_gauss_table(count, gt)
        N, dN = np.empty((count, 10)), np.empty((count, 3, 10))
        for g in range(count):
            lib.orc_shape_functions(gt[g, 1], gt[g, 2], gt[g, 3], N[g], dN[g])
        return gt, N, dN

    @staticmethod
    def model_eval(model, lam, mu, F):
        lib = PortOracle.lib()
        lib.orc_stress.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp]
        lib.orc_ctensor.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp]
        F = np.ascontiguousarray(F, np.float64)
        S, ct = np.empty((3, 3)), np.empty((3, 3, 3, 3))
        lib.orc_stress(model, lam, mu, F, S)
        lib.orc_ctensor(model, lam, mu, F, ct)
        return S, ct

    @staticmethod
    def matmul(which, A, B):
        lib = PortOracle.lib()
        f = [lib.orc_mul3, lib.orc_mul3_tn, lib.orc_mul3_nt][which]
        f.argtypes = [_dp, _dp, _dp]
        R = np.empty((3, 3))
        f(np.ascontiguousarray(A, np.float64), np.ascontiguousarray(B, np.float64), R)
        return R


# closed forms of exact-solutions/uniaxial (the reference's manual validation)


def uniaxial_neohookean(k1, lam=100.0, mu=100.0):
    """uniaxial_neohookean_bonet.m:20-37: solve mu(k2^2-1)+lam ln(k1 k2^2)=0, then sigma."""
    k2 = 1.0
    for _ in range(100):
        J = k1 * k2 * k2
        f = mu * (k2 * k2 - 1) + lam * np.log(J)
        df = 2 * mu * k2 + 2 * lam / k2
        step = f / df
        k2 -= step
        if abs(step) < 1e-16:
            break
    J = k1 * k2 * k2
    return k2, (mu * (k1 * k1 - 1) + lam * np.log(J)) / J


def uniaxial_a5(k1, lam=100.0, mu=100.0):
    """uniaxial.m:20,35 with n = 5."""
    kk1 = k1 ** 2
    kk2 = (3 * lam + 2 * mu - lam * kk1) / (2.0 * lam + 2.0 * mu)
    k2 = kk2 ** 0.5
    s = (k1 / k2 ** 2) * ((lam + 2 * mu) * kk1 + 2 * lam * kk2 - (3 * lam + 2 * mu)) / 2.0
    return k2, s
