"""ctypes binding of include/fea_gpu.h (lib/libfea_gpu.so).

Thin by design: numpy arrays in, numpy arrays out, every call goes through the C-ABI a
C host would use.  There is no Python or CPU implementation behind these methods --
if the shared library (or a CUDA device) is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FEA_GPU_LIB: another build of the same library (A/B timing of kernel variants from tools/)
LIB_PATH = os.environ.get("FEA_GPU_LIB") or os.path.normpath(os.path.join(_HERE, "..", "lib", "libfea_gpu.so"))

MODEL_A5, MODEL_NH = 0, 1
X0_ZERO, X0_RHS, ABS_TOL, ACCEPT_STALL = 0, 1, 2, 4
OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_MESH, ERR_NOT_CONVERGED, ERR_STALLED = range(7)

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

# every symbol include/fea_gpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fea_gpu_create", "fea_gpu_create_multi", "fea_gpu_destroy", "fea_gpu_nccl_unique_id", "fea_gpu_last_error",
    "fea_gpu_set_nodes", "fea_gpu_get_nodes", "fea_gpu_apply_increment", "fea_gpu_update_nodes",
    "fea_gpu_update_nodes_scaled", "fea_gpu_save_nodes", "fea_gpu_restore_nodes", "fea_gpu_extrapolate_nodes",
    "fea_gpu_update_state", "fea_gpu_assemble_stiffness", "fea_gpu_assemble_residual",
    "fea_gpu_assemble_all", "fea_gpu_apply_bc", "fea_gpu_save_stiffness", "fea_gpu_restore_stiffness",
    "fea_gpu_solve", "fea_gpu_dot_R_u", "fea_gpu_spmv", "fea_gpu_get_state", "fea_gpu_get_forces",
    "fea_gpu_set_forces", "fea_gpu_get_solution", "fea_gpu_get_csr", "fea_gpu_get_element_matrix", "fea_gpu_get_state_elems", "fea_gpu_bad_points",
    "fea_gpu_counts", "fea_gpu_launch_count", "fea_gpu_timer_start", "fea_gpu_timer_stop",
    "fea_gpu_sync", "fea_gpu_phase_ms", "fea_gpu_bench_spmv", "fea_gpu_measure_peaks",
    "fea_gpu_flush_l2", "fea_gpu_set_param", "fea_gpu_measure_dmma", "fea_gpu_bench_comm", "fea_gpu_host_alloc", "fea_gpu_host_free", "fea_gpu_step_from_host", "fea_plan_create", "fea_plan_destroy", "fea_plan_counts", "fea_plan_arrays",
    "fea_plan_node_owner", "fea_plan_sell_arrays", "fea_plan_cell_arrays", "fea_mesh_block", "fea_mesh_cylinder",
]

_lib = None


class FeaGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fea_gpu error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FeaGpuError(ERR_CUDA, f"{LIB_PATH} not built (run `make -C fea-large_b200 lib`); no CPU fallback exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.fea_gpu_last_error.restype = C.c_char_p
        _lib.fea_gpu_launch_count.restype = C.c_int64
    return _lib


def _check(rc, allow=()):
    if rc != OK and rc not in allow:
        raise FeaGpuError(rc, lib().fea_gpu_last_error().decode())
    return rc


def launch_count() -> int:
    return int(lib().fea_gpu_launch_count())


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().fea_gpu_nccl_unique_id(buf))
    return buf.raw


def measure_peaks(device=0):
    a, b = C.c_double(0), C.c_double(0)
    _check(lib().fea_gpu_measure_peaks(int(device), C.byref(a), C.byref(b)))
    return a.value, b.value


def measure_dmma(device=0) -> float:
    a = C.c_double(0)
    lib().fea_gpu_measure_dmma.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    _check(lib().fea_gpu_measure_dmma(int(device), C.byref(a)))
    return a.value


def mesh_block(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, y0=0.0, bc_style=0, dy=0.0):
    """Kuhn 6-tet block of 10-node tets (see fea_mesh_block).  Returns dict of arrays."""
    f = lib().fea_mesh_block
    f.argtypes = [C.c_int32] * 3 + [C.c_double] * 4 + [C.c_int32, C.c_double] + [C.c_void_p] * 8
    nn, ne, npz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _check(f(nx, ny, nz, lx, ly, lz, y0, bc_style, dy, C.addressof(nn), C.addressof(ne), C.addressof(npz),
             None, None, None, None, None))
    nodes = np.empty((nn.value, 3))
    conn = np.empty((ne.value, 10), np.int32)
    pn = np.empty(npz.value, np.int32)
    pt = np.empty(npz.value, np.int32)
    pv = np.empty((npz.value, 3))
    _check(f(nx, ny, nz, lx, ly, lz, y0, bc_style, dy, None, None, None, nodes.ctypes.data, conn.ctypes.data,
             pn.ctypes.data, pt.ctypes.data, pv.ctypes.data))
    return dict(nodes=nodes, conn=conn, presc_node=pn, presc_type=pt, presc_vals=pv)


def mesh_cylinder(nr, nt, nz, r_in=1.0, r_out=2.0, length=1.0, delta=0.0):
    """Hollow cylinder about z with prescribed radial wall displacement (see fea_mesh_cylinder)."""
    f = lib().fea_mesh_cylinder
    f.argtypes = [C.c_int32] * 3 + [C.c_double] * 4 + [C.c_void_p] * 8
    nn, ne, npz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _check(f(nr, nt, nz, r_in, r_out, length, delta, C.addressof(nn), C.addressof(ne), C.addressof(npz),
             None, None, None, None, None))
    nodes = np.empty((nn.value, 3))
    conn = np.empty((ne.value, 10), np.int32)
    pn = np.empty(npz.value, np.int32)
    pt = np.empty(npz.value, np.int32)
    pv = np.empty((npz.value, 3))
    _check(f(nr, nt, nz, r_in, r_out, length, delta, None, None, None, nodes.ctypes.data, conn.ctypes.data,
             pn.ctypes.data, pt.ctypes.data, pv.ctypes.data))
    return dict(nodes=nodes, conn=conn, presc_node=pn, presc_type=pt, presc_vals=pv)


def host_array(shape, dtype=np.float64):
    """numpy array backed by page-locked memory from fea_gpu_host_alloc (never freed: bench use)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    lib().fea_gpu_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    _check(lib().fea_gpu_host_alloc(C.byref(p), n))
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Plan:
    """Host-only partition + symbolic phase (no CUDA)."""

    def __init__(self, nodes, conn, rank=0, nranks=1):
        nodes = np.ascontiguousarray(nodes, np.float64)
        conn = np.ascontiguousarray(conn, np.int32)
        self.h = C.c_void_p()
        f = lib().fea_plan_create
        f.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _dp, _ip, C.c_int32, C.c_int32]
        _check(f(C.byref(self.h), len(nodes), len(conn), nodes, conn, rank, nranks))
        self.n_nodes = len(nodes)
        cnt = np.zeros(16, np.int64)
        lib().fea_plan_counts.argtypes = [C.c_void_p, _lp]
        _check(lib().fea_plan_counts(self.h, cnt))
        self.counts = cnt
        (self.n_own, self.n_local, self.n_elems, self.nnzb, self.n_contrib, self.n_nbr,
         self.n_send, self.n_ghost) = (int(v) for v in cnt[:8])
        g = lib().fea_plan_arrays
        g.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        self.node_gid = np.empty(self.n_local, np.int32)
        self.elem_gid = np.empty(self.n_elems, np.int32)
        self.browptr = np.empty(self.n_own + 1, np.int32)
        self.bcol = np.empty(self.nnzb, np.int32)
        self.cptr = np.empty(self.nnzb + 1, np.int32)
        self.csrc = np.empty(self.n_contrib, np.uint32)
        self.nbr_rank = np.empty(self.n_nbr, np.int32)
        self.send_ptr = np.empty(self.n_nbr + 1, np.int32)
        self.send_nodes = np.empty(self.n_send, np.int32)
        self.recv_ptr = np.empty(self.n_nbr + 1, np.int32)
        _check(g(self.h, *[a.ctypes.data for a in (self.node_gid, self.elem_gid, self.browptr, self.bcol,
                                                    self.cptr, self.csrc, self.nbr_rank, self.send_ptr,
                                                    self.send_nodes, self.recv_ptr)]))
        self.owner = np.empty(self.n_nodes, np.int32)
        lib().fea_plan_node_owner.argtypes = [C.c_void_p, _ip]
        _check(lib().fea_plan_node_owner(self.h, self.owner))
        self.n_slots, self.n_slices = int(cnt[10]), int(cnt[11])
        self.slice_ptr = np.empty(self.n_slices + 1, np.int32)
        self.sell_row = np.empty(32 * self.n_slices, np.int32)
        self.sbcol = np.empty(self.n_slots, np.int32)
        self.scptr = np.empty(self.n_slots + 1, np.int32)
        self.scsrc = np.empty(self.n_contrib, np.uint32)
        self.sdiag = np.empty(self.n_own, np.int32)
        f = lib().fea_plan_sell_arrays
        f.argtypes = [C.c_void_p] * 7
        _check(f(self.h, *[a.ctypes.data for a in (self.slice_ptr, self.sell_row, self.sbcol, self.scptr,
                                                    self.scsrc, self.sdiag)]))

        self.n_cells, self.n_cols_active = int(cnt[12]), int(cnt[13])
        ne_pad = (self.n_elems + 31) // 32 * 32
        self.cmeta = np.empty(self.n_slots, np.uint16)
        self.ccell = np.empty(self.n_slots // 32 + 1, np.int32)
        self.cmirror = np.empty(self.n_slots, np.int32)
        self.edest = np.empty((55, ne_pad), np.uint32)
        self.col_ready = np.empty(self.n_slots // 32, np.int32)
        self.col_order = np.empty(self.n_cols_active, np.int32)
        f = lib().fea_plan_cell_arrays
        f.argtypes = [C.c_void_p] * 7
        _check(f(self.h, *[a.ctypes.data for a in (self.cmeta, self.ccell, self.cmirror, self.edest, self.col_ready,
                                                    self.col_order)]))

    def close(self):
        if self.h:
            lib().fea_plan_destroy.argtypes = [C.c_void_p]
            lib().fea_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FeaGpu:
    """One GPU context (= one rank).  Methods mirror the C-ABI one to one."""

    def __init__(self, nodes, conn, model, lam, mu, n_gauss=5, presc_node=None, presc_type=None,
                 presc_vals=None, rank=0, nranks=1, nccl_id=None, device=0, n_gpus=None):
        L = lib()
        nodes = np.ascontiguousarray(nodes, np.float64)
        conn = np.ascontiguousarray(conn, np.int32)
        n_presc = 0 if presc_node is None else len(presc_node)
        pn = np.ascontiguousarray(presc_node if n_presc else np.zeros(1), np.int32)
        pt = np.ascontiguousarray(presc_type if n_presc else np.zeros(1), np.int32)
        pv = np.ascontiguousarray(presc_vals if n_presc else np.zeros((1, 3)), np.float64)
        self.n_nodes, self.n_elems, self.ng = len(nodes), len(conn), n_gauss
        self.n = 3 * self.n_nodes
        self.h = C.c_void_p()
        if n_gpus is not None:      # one process, several GPUs (fea_gpu_create_multi)
            f = L.fea_gpu_create_multi
            f.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _dp, _ip, C.c_int32, C.c_double, C.c_double,
                          C.c_int32, C.c_int32, _ip, _ip, _dp, C.c_int32, C.c_void_p]
            _check(f(C.byref(self.h), self.n_nodes, self.n_elems, nodes, conn, model, lam, mu, n_gauss, n_presc,
                     pn, pt, pv, int(n_gpus), None))
        else:
            f = L.fea_gpu_create
            f.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _dp, _ip, C.c_int32, C.c_double, C.c_double,
                          C.c_int32, C.c_int32, _ip, _ip, _dp, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
            _check(f(C.byref(self.h), self.n_nodes, self.n_elems, nodes, conn, model, lam, mu, n_gauss, n_presc,
                     pn, pt, pv, rank, nranks, nccl_id, device))
        for name in ("update_state", "assemble_stiffness", "assemble_residual", "update_nodes",
                     "save_stiffness", "restore_stiffness", "sync", "timer_start", "flush_l2", "save_nodes",
                     "restore_nodes"):
            getattr(L, "fea_gpu_" + name).argtypes = [C.c_void_p]

    def close(self):
        if getattr(self, "h", None):
            lib().fea_gpu_destroy.argtypes = [C.c_void_p]
            lib().fea_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- simple phase calls ----------------------------------------------------
    def _simple(self, name):
        _check(getattr(lib(), "fea_gpu_" + name)(self.h))

    def update_state(self): self._simple("update_state")
    def assemble_stiffness(self): self._simple("assemble_stiffness")
    def assemble_residual(self): self._simple("assemble_residual")
    def update_nodes(self): self._simple("update_nodes")
    def save_nodes(self): self._simple("save_nodes")
    def restore_nodes(self): self._simple("restore_nodes")

    def extrapolate_nodes(self, alpha=1.0):
        lib().fea_gpu_extrapolate_nodes.argtypes = [C.c_void_p, C.c_double]
        _check(lib().fea_gpu_extrapolate_nodes(self.h, float(alpha)))

    def update_nodes_scaled(self, eta):
        lib().fea_gpu_update_nodes_scaled.argtypes = [C.c_void_p, C.c_double]
        _check(lib().fea_gpu_update_nodes_scaled(self.h, float(eta)))

    def save_stiffness(self): self._simple("save_stiffness")
    def restore_stiffness(self): self._simple("restore_stiffness")
    def sync(self): self._simple("sync")
    def timer_start(self): self._simple("timer_start")
    def flush_l2(self): self._simple("flush_l2")

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        lib().fea_gpu_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _check(lib().fea_gpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def assemble_all(self, with_stiffness=True, fuse_bc=False):
        lib().fea_gpu_assemble_all.argtypes = [C.c_void_p, C.c_int32]
        _check(lib().fea_gpu_assemble_all(self.h, (1 if with_stiffness else 0) | (2 if fuse_bc else 0)))

    def apply_increment(self, lam=1.0):
        lib().fea_gpu_apply_increment.argtypes = [C.c_void_p, C.c_double]
        _check(lib().fea_gpu_apply_increment(self.h, float(lam)))

    def apply_bc(self, lam=0.0):
        lib().fea_gpu_apply_bc.argtypes = [C.c_void_p, C.c_double]
        _check(lib().fea_gpu_apply_bc(self.h, float(lam)))

    def solve(self, tol=1e-14, max_iter=20000, flags=X0_ZERO, allow_unconverged=False, accept_stall=False):
        """(iterations, relative residual, ok).  A solve that ends on the stall / divergence guard raises
        (FEA_GPU_ERR_STALLED) unless accept_stall; allow_unconverged returns ok = False instead of raising
        for both max_iter and the guard."""
        it, rr = C.c_int32(0), C.c_double(0)
        f = lib().fea_gpu_solve
        f.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        rc = _check(f(self.h, tol, max_iter, flags | (ACCEPT_STALL if accept_stall else 0), C.byref(it), C.byref(rr)),
                    allow=(ERR_NOT_CONVERGED, ERR_STALLED) if allow_unconverged else ())
        return it.value, rr.value, rc == OK

    def dot_R_u(self) -> float:
        out = C.c_double(0)
        lib().fea_gpu_dot_R_u.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _check(lib().fea_gpu_dot_R_u(self.h, C.byref(out)))
        return out.value

    # --- host <-> device vectors ----------------------------------------------
    def set_nodes(self, x):
        lib().fea_gpu_set_nodes.argtypes = [C.c_void_p, _dp]
        _check(lib().fea_gpu_set_nodes(self.h, np.ascontiguousarray(x, np.float64)))

    def _getvec(self, name, shape):
        out = np.zeros(shape)
        f = getattr(lib(), "fea_gpu_" + name)
        f.argtypes = [C.c_void_p, _dp]
        _check(f(self.h, out))
        return out

    def get_nodes(self): return self._getvec("get_nodes", (self.n_nodes, 3))
    def get_forces(self): return self._getvec("get_forces", self.n)
    def get_solution(self): return self._getvec("get_solution", self.n)

    def set_forces(self, r):
        lib().fea_gpu_set_forces.argtypes = [C.c_void_p, _dp]
        _check(lib().fea_gpu_set_forces(self.h, np.ascontiguousarray(r, np.float64)))

    def spmv(self, x):
        y = np.zeros(self.n)
        lib().fea_gpu_spmv.argtypes = [C.c_void_p, _dp, _dp]
        _check(lib().fea_gpu_spmv(self.h, np.ascontiguousarray(x, np.float64), y))
        return y

    def get_state(self):
        F = np.zeros((self.n_elems, self.ng, 3, 3))
        S = np.zeros((self.n_elems, self.ng, 3, 3))
        lib().fea_gpu_get_state.argtypes = [C.c_void_p, _dp, _dp]
        _check(lib().fea_gpu_get_state(self.h, F, S))
        return F, S

    def get_state_elems(self, elems):
        """F, S [n][ng][3][3] and a mask of the listed (global) elements that are local to this rank."""
        elems = np.ascontiguousarray(elems, np.int32)
        F = np.zeros((len(elems), self.ng, 3, 3))
        S = np.zeros((len(elems), self.ng, 3, 3))
        found = np.zeros(len(elems), np.int32)
        lib().fea_gpu_get_state_elems.argtypes = [C.c_void_p, C.c_int32, _ip, _dp, _dp, _ip]
        _check(lib().fea_gpu_get_state_elems(self.h, len(elems), elems, F, S, found))
        return F, S, found.astype(bool)

    def get_csr(self):
        f = lib().fea_gpu_get_csr
        f.argtypes = [C.c_void_p] * 7
        nr, nz = C.c_int64(0), C.c_int64(0)
        _check(f(self.h, C.addressof(nr), C.addressof(nz), None, None, None, None))
        rows = np.empty(nr.value, np.int32)
        rp = np.empty(nr.value + 1, np.int32)
        ci = np.empty(nz.value, np.int32)
        v = np.empty(nz.value)
        _check(f(self.h, None, None, rows.ctypes.data, rp.ctypes.data, ci.ctypes.data, v.ctypes.data))
        return rows, rp, ci, v

    def element_matrix(self, e):
        """Dense 30x30 K_e of global element e from the staging buffer; None if e is not on this rank."""
        ke = np.zeros((30, 30))
        lib().fea_gpu_get_element_matrix.argtypes = [C.c_void_p, C.c_int32, _dp]
        rc = lib().fea_gpu_get_element_matrix(self.h, int(e), ke)
        if rc == ERR_ARG:
            return None
        _check(rc)
        return ke

    def bad_points(self) -> int:
        out = C.c_int64(0)
        lib().fea_gpu_bad_points.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        _check(lib().fea_gpu_bad_points(self.h, C.byref(out)))
        return out.value

    def counts(self):
        out = np.zeros(16, np.int64)
        lib().fea_gpu_counts.argtypes = [C.c_void_p, _lp]
        _check(lib().fea_gpu_counts(self.h, out))
        keys = ["owned_nodes", "local_nodes", "local_elems", "nnzb", "contribs", "neighbours",
                "halo_sent", "halo_recv", "global_nodes", "global_elems", "sell_slots", "sell_slices", "gather9"]
        return dict(zip(keys, (int(v) for v in out)))

    def phase_ms(self):
        out = np.zeros(16)
        lib().fea_gpu_phase_ms.argtypes = [C.c_void_p, _dp]
        _check(lib().fea_gpu_phase_ms(self.h, out))
        keys = ["element", "gather_k", "gather_r", "bc", "pcg", "spmv_avg", "halo"]
        d = dict(zip(keys, (float(v) for v in out)))
        d["phase_samples"] = int(out[14])
        d["spmv_samples"] = int(out[8])
        d["pcg_iters"] = int(out[9])
        d["pcg_exit"] = int(out[10])
        d["pcg_best_relres"], d["pcg_last_relres"], d["pcg_stall"] = float(out[11]), float(out[12]), int(out[13])
        return d

    def set_param(self, name, value):
        lib().fea_gpu_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        _check(lib().fea_gpu_set_param(self.h, name.encode(), float(value)))

    def step_from_host(self, x, R, with_stiffness=True):
        """fea_gpu_step_from_host: x, R are host arrays (ideally from host_array()); returns bytes moved."""
        f = lib().fea_gpu_step_from_host
        f.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(f(self.h, x, int(with_stiffness), R, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bench_comm(self, reps=50):
        """(halo ms, all-reduce ms), each collective timed alone; zeros on one rank."""
        a, b = C.c_double(0), C.c_double(0)
        lib().fea_gpu_bench_comm.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _check(lib().fea_gpu_bench_comm(self.h, reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bench_spmv(self, reps=20) -> float:
        ms = C.c_double(0)
        lib().fea_gpu_bench_spmv.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
        _check(lib().fea_gpu_bench_spmv(self.h, reps, C.byref(ms)))
        return ms.value
