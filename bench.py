#!/usr/bin/env python
"""Benchmark of the fea-large hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]

A *step* is one pass of the element hot path over the whole mesh: per-element geometry,
F, stress and tangent, K_e and R_e, deterministic gather into the global block-CSR matrix
and residual, Dirichlet cancellation (the reference's solver_create_current_shape_gradients
+ _stresses + _stiffness + _residual_forces + solver_apply_prescribed_bc, fea_solver.c:
171-203).  `value` = element assemblies per second, whole job, inputs resident in HBM;
`e2e` = the same pass through fea_gpu_step_from_host with host nodes in / host residual
out.  The Newton-iteration side of BASELINE.json's metric (assembly + Jacobi-PCG solve to
the reference's tolerance + update) is measured in the same run, outside the K timed
steps, and reported under "newton"; "strong_c4" repeats it on the fixed 10.3 M-DOF
Neo-Hookean cube of BASELINE configs[3] partitioned over the N ranks.

Before anything is timed the run checks itself against the CPU oracle ("parity"): a window of
the bench mesh in the bench state (F, sigma, K_e, K through probe vectors, R) on every rank,
and for N > 1 a whole Newton step on N ranks against one rank and the oracle.  A failed check
makes the process exit non-zero.

Workload at N GPUs: Kuhn block of n x (n*N) x n cubes of unit size (weak scaling, n = 55
-> 998 250 ten-node tets and 4 102 893 DOF per GPU = BASELINE.json configs[2]), A5
compressible, lambda = mu = 100, 5-point rule, deformed by the exact uniaxial map at
stretch 1.5 plus a seeded perturbation of 1e-3 h.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ELEM = 29000.0     # index form, 5-point rule, FMA = 2
BYTES_PER_ELEM = 3600.0     # compulsory HBM traffic of a fused assembly
METRIC = "element_assemblies_per_sec"
PARITY_TOL_ELEM = 1e-10     # F, sigma, K_e, K probes, R against the oracle, relative to the largest entry.
                            # (1e-12 on the shipped bricks; F = (sum grad N (x) X)^-1 with absolute coordinates,
                            # fea_solver.c:1141-1152, is conditioned like |X| / h, up to ~900 on the 8-GPU bar)
PARITY_TOL_U = 1e-8         # displacements after a linear solve / Newton step (1e-9 in the tests)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL's version banner, for one) write to stdout; the contract is ONE JSON line
# there.  Keep the real stdout aside and point fd 1 at stderr for everything else.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def kernel_source_hash():
    """sha1 of the CUDA sources: profiles/ncu_traffic.json is only quoted while it describes them."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "fea-large_b200", "csrc")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)             # let the first samples arrive before the timed region starts
            self.n_before = len(self.rows)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_end(self):
        self.n_end = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        self.t.join(timeout=2)
        rows = self.rows[max(getattr(self, "n_before", 1) - 1, 0):getattr(self, "n_end", len(self.rows)) + 2]
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def lateral_stretch(k1, model, lam=100.0, mu=100.0):
    """Lateral stretch of the homogeneous uniaxial state (closed forms behind
    exact-solutions/uniaxial); only used to build a physically sensible benchmark state."""
    if model == 0:   # A5: k2^2 = (3 lam + 2 mu - lam k1^2) / (2 lam + 2 mu)
        return float(np.sqrt((3 * lam + 2 * mu - lam * k1 * k1) / (2 * lam + 2 * mu)))
    k2 = 1.0         # NH: mu (k2^2 - 1) + lam ln(k1 k2^2) = 0
    for _ in range(60):
        f = mu * (k2 * k2 - 1) + lam * np.log(k1 * k2 * k2)
        k2 -= f / (2 * mu * k2 + 2 * lam / k2)
    return float(k2)


def perturbation(node_ids, h, seed=12345):
    """Uniform perturbation in [-1e-3 h, 1e-3 h)^3 as a pure function of the GLOBAL node id, so a window
    of the mesh can be put into the bench state without generating the state of the whole mesh."""
    ids = np.asarray(node_ids, np.uint64)
    out = np.empty((len(ids), 3))
    for d in range(3):       # splitmix64 of (3 id + d + seed)
        z = (ids * np.uint64(3) + np.uint64(d) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        out[:, d] = (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    return (out - 0.5) * 2e-3 * h


def deformed_state(nodes, h, seed=12345, k1=1.5, model=0, node_ids=None):
    """Exact uniaxial map at stretch k1 plus a uniform perturbation of amplitude 1e-3 h."""
    k2 = lateral_stretch(k1, model)
    ids = np.arange(len(nodes)) if node_ids is None else node_ids
    with np.errstate(over="ignore"):
        return nodes * np.array([k2, k1, k2]) + perturbation(ids, h, seed)


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own compiled element code (oracle/_ref) or
# the plain-C port, on a bounded sample of the same workload.  Nothing here loads the product
# library: the sample mesh comes from the numpy mesher under oracle/.


def _oracle_block(n_s, model, seed):
    from oracle.kuhn import kuhn_block
    from oracle.oracle import Model, PortOracle, RefOracle, have_ref
    mb = kuhn_block(n_s, n_s, n_s, float(n_s), float(n_s), float(n_s), 0.0, 1, 0.0)
    m = Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=mb["presc_node"], presc_type=mb["presc_type"],
              presc_vals=mb["presc_vals"], model=model, lam=100.0, mu=100.0, gauss=5)
    kind = "reference" if have_ref() else "port"
    o = RefOracle(m) if kind == "reference" else PortOracle(m)
    o.set_nodes(deformed_state(m.nodes, 0.5, seed, model=model))
    return m, o, kind


def _cpu_sample_worker(args):
    n_s, model, steps, warmup, seed = args
    sys.path.insert(0, ROOT)
    m, o, kind = _oracle_block(n_s, model, seed)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        o.update_state()            # solver_create_current_shape_gradients + _stresses
        o.assemble_stiffness()      # solver_create_stiffness (9000 sp_matrix_element_add / element)
        o.assemble_residual()       # solver_create_residual_forces
        o.apply_bc(0.0)             # solver_apply_prescribed_bc(0)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return len(m.conn), times, kind


def cpu_sample(n_s, model, steps, warmup, workers):
    if workers == 1:
        res = [_cpu_sample_worker((n_s, model, steps, warmup, 12345))]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_sample_worker, [(n_s, model, steps, warmup, 12345 + w) for w in range(workers)])
    ne = res[0][0]
    wall = max(sum(r[1]) for r in res)           # slowest worker bounds the throughput
    value = workers * ne * steps / wall
    return value, wall / steps, ne, res[0][2]


def cpu_newton_iteration(n_s, model):
    """One Newton iteration of the reference's loop (fea_solver.c:180-221) on one host core: increment,
    state, stiffness, residual, Dirichlet cancellation, linear solve, update."""
    from oracle.kuhn import kuhn_block
    from oracle.oracle import Model, PortOracle, RefOracle, have_ref
    mb = kuhn_block(n_s, n_s, n_s, float(n_s), float(n_s), float(n_s), 0.0, 1, 0.005)
    m = Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=mb["presc_node"], presc_type=mb["presc_type"],
              presc_vals=mb["presc_vals"], model=model, lam=100.0, mu=100.0, gauss=5)
    kind = "reference" if have_ref() else "port"
    o = RefOracle(m) if kind == "reference" else PortOracle(m)
    t0 = time.perf_counter()
    o.apply_increment(1.0); o.update_state(); o.assemble_stiffness(); o.assemble_residual(); o.apply_bc(0.0)
    t1 = time.perf_counter()
    its = o.solve_slae()
    t2 = time.perf_counter()
    o.update_with_solution(); o.update_state()
    t3 = time.perf_counter()
    return {"newton_iters_per_sec": 1.0 / (t3 - t0), "seconds": t3 - t0, "assembly_seconds": t1 - t0,
            "solve_seconds": t2 - t1, "pcg_iters": int(its), "elements": len(m.conn), "dof": int(m.n_dof),
            "cores": 1, "kind": kind,
            "linear_solve": "Jacobi-PCG stand-in for libspmatrix (absent from the reference tree), stops at relative "
                            "residual 1e-15 or its rounding floor (the GPU arm stops at 1e-14)",
            "sample": f"Kuhn block of {n_s}^3 cubes, first iteration of a 0.01 h increment"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same pass on the box's
    host cores (all of them: one independent sample per core, the reference itself is
    single-threaded by construction)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_s = args.ref_sample
    value, sec_per_step, ne, kind = cpu_sample(n_s, args.model, args.steps, max(args.warmup, 1), cores)
    sample = (f"bounded sample: {cores} independent Kuhn sub-blocks of {n_s}^3 cubes ({ne} tets each, {cores * ne} in all), "
              f"same state and pass as the GPU arm; reference objects compiled from /root/reference with a sorted-array "
              f"stand-in for libspmatrix"
              if kind == "reference" else f"bounded sample: {cores} x {ne} tets, plain-C port of the reference (oracle/oracle_fea.c)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, world):
    n = args.n
    return {"workload": f"Kuhn block {n}x{n * world}x{n} cubes, 10-node tets, "
                        f"{'A5' if args.model == 0 else 'Neo-Hookean'} compressible, lambda=mu=100, 5-pt rule "
                        f"(BASELINE configs[2] per GPU)",
            "elements": 6 * n * n * n * world, "dof": 3 * (2 * n + 1) * (2 * n * world + 1) * (2 * n + 1),
            "parallelism": f"row/element partition over {world} GPU(s), recursive coordinate bisection (slabs along y for this bar)",
            "l2": "working set (K_e staging + matrix values, >7 GB per GPU) exceeds L2; no flush needed",
            "state": "uniaxial map at stretch 1.5 + 1e-3 h perturbation, seed 12345",
            "bc": "faces y=min / y=max clamped (type 7, as data/*_brick.sexp), 0.01 h per increment"}


# ---------------------------------------------------------------------------------------
# parity: the run checks itself against the CPU oracle before anything is timed


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def window_parity(g, args, world, rank, n_nodes_full):
    """A window of ns^3 cubes of the bench mesh, in the bench state, through the reference-compiled
    element code (oracle/_ref; the plain-C port where that is absent) against what this rank's GPU holds
    after one assembly pass: F and sigma of the window's elements, K_e of seeded elements (capture at the
    reference's sp_matrix_element_add call sites), and -- for the window's interior nodes, whose rows are
    complete inside the window -- the assembled K through probe vectors and the residual rows.
    For N > 1 the window straddles the rank 0 / rank 1 interface.  Collective (spmv / get_forces are)."""
    from oracle.kuhn import kuhn_block
    from oracle.oracle import Model, PortOracle, RefOracle, have_ref
    n, ns = args.n, args.parity_sample
    full = (n, n * world, n)
    oy = (n - ns) // 2 if world == 1 else n - ns // 2
    ox = oz = (n - ns) // 2
    w = kuhn_block(ns, ns, ns, float(n), float(n * world), float(n), 0.0, cube_origin=(ox, oy, oz), full=full)
    m = Model(nodes=w["nodes"], conn=w["conn"], presc_node=w["presc_node"], presc_type=w["presc_type"],
              presc_vals=w["presc_vals"], model=args.model, lam=100.0, mu=100.0, gauss=5)
    kind = "reference" if have_ref() else "port"
    o = RefOracle(m) if kind == "reference" else PortOracle(m)
    xw = deformed_state(w["nodes"], 0.5, model=args.model, node_ids=w["node_gid"])
    o.set_nodes(xw); o.update_state(); o.assemble_stiffness(); o.assemble_residual()
    Fo, So = o.get_state()
    Ro = o.get_forces()
    rpo, cio, vo = o.get_csr()
    # interior half-grid nodes of the window: every element they belong to lies in the window
    p = 2 * ns + 1
    jy, jz, jx = np.meshgrid(np.arange(p), np.arange(p), np.arange(p), indexing="ij")
    interior = ((jx > 0) & (jx < p - 1) & (jy > 0) & (jy < p - 1) & (jz > 0) & (jz < p - 1)).reshape(-1)
    gid = w["node_gid"]
    idof_w = (3 * np.nonzero(interior)[0][:, None] + np.arange(3)).reshape(-1)          # window dof ids
    idof_g = (3 * gid[interior][:, None] + np.arange(3)).reshape(-1)                     # global dof ids
    err = {}
    # GPU side: one plain assembly pass in the bench state (no Dirichlet folding: the window has none)
    g.assemble_all(True)
    Fg, Sg, found = g.get_state_elems(w["elem_gid"])
    err["F"] = relmax(Fg[found], Fo[found])
    err["sigma"] = relmax(Sg[found], So[found])
    rng = np.random.default_rng(777)
    picks = rng.choice(len(w["elem_gid"]), size=min(args.parity_elems, len(w["elem_gid"])), replace=False)
    ke_err, ke_n = 0.0, 0
    for k in picks:
        ke = g.element_matrix(int(w["elem_gid"][k]))
        if ke is None:
            continue
        ke_err = max(ke_err, relmax(ke, o.element_matrix(int(k))))
        ke_n += 1
    err["K_e"] = ke_err
    Rg = g.get_forces()
    err["R_rows"] = relmax(Rg[idof_g], Ro[idof_w])
    kerr = 0.0
    for s in range(2):
        xv = np.zeros(3 * n_nodes_full)
        pv = rng.standard_normal(len(idof_g))
        xv[idof_g] = pv
        yg = g.spmv(xv)
        xo = np.zeros(m.n_dof)
        xo[idof_w] = pv
        yo = np.zeros(m.n_dof)
        np.add.at(yo, np.repeat(np.arange(m.n_dof), np.diff(rpo)), vo * xo[cio])
        dof_all_g = (3 * gid[:, None] + np.arange(3)).reshape(-1)
        kerr = max(kerr, relmax(yg[dof_all_g], yo))
        outside = np.ones(len(yg), bool)
        outside[dof_all_g] = False
        kerr = max(kerr, float(np.abs(yg[outside]).max()) / max(np.abs(yo).max(), 1e-300))   # nothing leaks out of the window
    err["K_probes"] = kerr
    info = {"oracle": kind, "window_cubes": ns, "window_elements": int(len(w["elem_gid"])),
            "elements_on_this_rank": int(found.sum()), "K_e_compared_on_this_rank": ke_n,
            "window_origin_cubes": [ox, oy, oz], "interior_rows": int(len(idof_g))}
    return err, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 60: ~0.2 s of device time, enough clock samples; 20 for --impl reference, whose steps are seconds of CPU work)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--n", type=int, default=55, help="cubes per edge per GPU (55 -> 998 250 tets)")
    ap.add_argument("--model", type=int, default=0, help="0 = A5 (configs[2]), 1 = Neo-Hookean")
    ap.add_argument("--newton-iters", type=int, default=2)
    ap.add_argument("--lin-tol", type=float, default=1e-14, help="PCG relative tolerance (reference files: 1e-14)")
    ap.add_argument("--lin-max-iter", type=int, default=20000)
    ap.add_argument("--ref-sample", type=int, default=11, help="cubes per edge of each CPU sample block (reference arm)")
    ap.add_argument("--cpu-baseline-sample", type=int, default=14)
    ap.add_argument("--cpu-newton-sample", type=int, default=10)
    ap.add_argument("--parity-sample", type=int, default=8, help="cubes per edge of the oracle-checked window")
    ap.add_argument("--parity-elems", type=int, default=64, help="seeded elements whose K_e is compared")
    ap.add_argument("--c4-n", type=int, default=75, help="cubes per edge of the strong-scaling cube (75 -> 10.3 M DOF)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 60
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return

    sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import fea_gpu as fg

    dist = None
    if world > 1:
        import torch.distributed as dist   # host-side plumbing only (gloo): id broadcast, barrier, max
        dist.init_process_group("gloo", rank=rank, world_size=world)

    def new_nccl_id():
        if dist is None:
            return None
        box = [fg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def barrier():
        if dist is not None:
            dist.barrier()

    def allred(v, op):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])

    def allmax(v): return allred(v, dist.ReduceOp.MAX) if dist is not None else v
    def allsum(v): return allred(v, dist.ReduceOp.SUM) if dist is not None else v
    def allmin(v): return allred(v, dist.ReduceOp.MIN) if dist is not None else v

    # ---- N > 1: a whole Newton step on N ranks against one rank and the oracle --------------------
    parity = {"ok": True, "n_ranks": world, "tol_elem": PARITY_TOL_ELEM, "tol_u": PARITY_TOL_U, "max_rel_elem": None,
              "max_rel_u": None}
    if world > 1 and not args.no_parity:
        from multirank_worker import newton_step_check
        ok_mr, e_mr = newton_step_check(dist, rank, world, local_rank, log=log)
        ok_mr = allmin(1.0 if ok_mr else 0.0) > 0.5
        parity["multirank_newton_step"] = {"ok": ok_mr, "errors": e_mr,
                                           "what": "4 x (4N+1) x 4 Neo-Hookean bar: R, u, <R,u>, x, sigma, host-buffer path on N ranks "
                                                   "vs one rank and the CPU oracle; single-reduction PCG variants where they converge (negative count = ended on its guard); bit-reproducible solve"}
        parity["ok"] = parity["ok"] and ok_mr

    n = args.n
    t0 = time.time()
    mb = fg.mesh_block(n, n * world, n, float(n), float(n * world), float(n), 0.0, 1, 0.01)
    nodes, conn = mb["nodes"], mb["conn"]
    n_elems, n_dof = len(conn), 3 * len(nodes)
    x_def = deformed_state(nodes, 0.5, model=args.model)
    g = fg.FeaGpu(nodes, conn, args.model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"],
                  rank=rank, nranks=world, nccl_id=new_nccl_id(), device=local_rank)
    cnt = g.counts()
    setup_s = allmax(time.time() - t0)
    if rank == 0:
        log(f"[bench] mesh+plan+upload {setup_s:.1f}s: {n_elems} tets, {n_dof} DOF, rank0 {cnt}")
    g.set_nodes(x_def)

    # ---- parity of the bench mesh itself, in the bench state --------------------------------------
    if not args.no_parity:
        tp = time.time()
        err, info = window_parity(g, args, world, rank, len(nodes))
        emax = {k: allmax(v) for k, v in err.items()}
        info["elements_checked_all_ranks"] = int(allsum(float(info["elements_on_this_rank"])))
        info["K_e_compared_all_ranks"] = int(allsum(float(info.pop("K_e_compared_on_this_rank"))))
        info.pop("elements_on_this_rank")
        max_elem = max(emax.values())
        parity["bench_mesh_window"] = {"errors": emax, **info, "ok": bool(max_elem <= PARITY_TOL_ELEM and info["K_e_compared_all_ranks"] > 0)}
        parity["max_rel_elem"] = max_elem
        parity["ok"] = parity["ok"] and parity["bench_mesh_window"]["ok"]
        if world > 1:
            e = parity["multirank_newton_step"]["errors"]
            parity["max_rel_u"] = max([v for k, v in e.items() if k in ("u", "u_oracle", "u_single_reduction", "x1")] or [0.0])
            parity["ok"] = parity["ok"] and parity["max_rel_u"] <= PARITY_TOL_U
            parity["max_rel_elem"] = max(max_elem, max([v for k, v in e.items() if k in ("R0", "R0_oracle")] or [0.0]))
        if rank == 0:
            log(f"[bench] parity ({time.time() - tp:.1f}s): {json.dumps(parity)}")

    def step():
        g.update_nodes()            # x += u (u = 0 here) and, for N > 1, the halo exchange of x
        g.assemble_all(True, fuse_bc=True)   # geometry, F, stress, tangent, K_e, R_e, both gathers with the
                                             # Dirichlet cancellation of solver_apply_prescribed_bc(0) folded in

    g.set_nodes(x_def)
    for _ in range(args.warmup):
        step()
    g.sync()
    g.phase_ms()                    # restart the per-kernel event averages: only timed steps count below
    launches0 = fg.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    g.sync()
    g.timer_start()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = g.timer_stop()             # CUDA events on the launching stream, synchronises
    barrier()
    wall = time.perf_counter() - wall0
    launches = fg.launch_count() - launches0
    ph = g.phase_ms()
    ms = allmax(ms)
    value = n_elems * args.steps / (ms * 1e-3)
    bad = g.bad_points()

    # per-kernel durations: CUDA events around every launch of the timed steps, averaged (no sync inside the loop)
    elem_ms, gather_ms, gres_ms, bc_ms, halo_x_ms = ph["element"], ph["gather_k"], ph["gather_r"], ph["bc"], ph["halo"]

    # ---- end to end: host nodes in, host residual out, every step ---------------------------
    xh = fg.host_array(nodes.shape)
    Rh = fg.host_array(n_dof)
    xh[:] = x_def
    for _ in range(2):
        g.step_from_host(xh, Rh)
    barrier()
    te = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h = g.step_from_host(xh, Rh)
    e2e_s = allmax(time.perf_counter() - te)
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None      # covers the device-resident and the end-to-end loops
    e2e_value = n_elems * args.steps / e2e_s
    h2d, d2h = allsum(float(h2d)), allsum(float(d2h))
    r_max = float(np.abs(Rh).max())

    # ---- Newton iterations: assembly + PCG to the reference's tolerance + update -------------
    halo_ms, allreduce_ms = g.bench_comm(50)
    newton = None
    if not args.no_newton:
        g.set_nodes(nodes)
        its, nt_ms, relres, spmv_ms, exits, pcg_ms = [], [], [], [], [], []
        for k in range(args.newton_iters):
            barrier()
            g.sync()
            g.timer_start()
            if k == 0:
                g.apply_increment(1.0)
            g.assemble_all(True)
            g.apply_bc(0.0)
            it, rr, ok = g.solve(args.lin_tol, args.lin_max_iter, fg.X0_ZERO, allow_unconverged=True)
            tol = g.dot_R_u()
            g.update_nodes()
            t_ms = allmax(g.timer_stop())
            p = g.phase_ms()
            its.append(it); nt_ms.append(t_ms); relres.append(rr); spmv_ms.append(p["spmv_avg"]); exits.append(p["pcg_exit"])
            pcg_ms.append(allmax(p["pcg"]))
            if rank == 0:
                log(f"[bench] newton it {k}: {t_ms:.1f} ms, pcg {it} its (relres {rr:.2e}, ok={ok}), <R,u>={tol:.3e}, "
                    f"spmv {p['spmv_avg']:.3f} ms")
        nnzb = allsum(float(cnt["nnzb"]))
        sp = float(np.mean(spmv_ms))
        newton = {"newton_iters_per_sec": len(nt_ms) / (sum(nt_ms) * 1e-3), "ms_per_newton_iter": float(np.mean(nt_ms)),
                  "pcg_iters_per_newton_iter": float(np.mean(its)), "pcg_relres": float(max(relres)),
                  "pcg_tol": args.lin_tol, "pcg_exit": exits, "pcg_exit_legend": "1 = tolerance met, 2 = stall/divergence guard, 0 = max_iter",
                  "pcg_iters_per_sec": float(sum(its) / (sum(nt_ms) * 1e-3)),
                  "pcg_ms_per_iter": float(sum(pcg_ms) / max(sum(its), 1)),
                  "pcg_variant": "classic PCG: all-reduce of p.Ap, then of (r.z, r.r), every iteration" if world > 1 else "classic PCG",
                  "spmv_ms": sp, "spmv_format": "3x3 blocks in SELL-32-sigma, fp64 values, int32 block columns",
                  "nnz_scalar_total": 9 * nnzb}
    comm = {"halo_ms": allmax(halo_ms), "allreduce_ms": allmax(allreduce_ms), "halo_x_ms_in_step": allmax(halo_x_ms),
            "how": "each collective timed alone, 50 back to back on the solve stream (CUDA events), max over ranks; "
                   "halo_x_ms_in_step = the exchange of x inside the timed assembly steps",
            "local_elements_max": int(allmax(float(cnt["local_elems"]))), "unique_elements_per_rank": n_elems // world,
            "halo_nodes_recv_max": int(allmax(float(cnt["halo_recv"]))), "setup_s": setup_s}
    peaks, peaks_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dfma_tf, copy_gbs = fg.measure_peaks(local_rank)
    dmma_tf = fg.measure_dmma(local_rank)
    g.close()

    # ---- strong scaling: the fixed 10.3 M-DOF Neo-Hookean cube of configs[3] over the N ranks ------
    strong = None
    if not args.no_c4:
        strong = strong_c4(args, fg, rank, world, local_rank, new_nccl_id, barrier, allmax, allsum)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        sys.exit(0 if parity["ok"] else 1)

    # DRAM traffic per launch from the committed ncu --set full capture: quoted only while the capture
    # describes the kernels that ran (same sources, same workload)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if (t["workload"]["n"] == args.n and t["workload"]["n_gpus"] == world and args.model == 0
                and t.get("csrc_sha1") == kernel_source_hash()):
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass

    # roofline of the assembly = the kernels of a timed step, against its compulsory HBM bytes
    asm_ms = elem_ms + gather_ms + gres_ms
    ach = BYTES_PER_ELEM * cnt["local_elems"] / (asm_ms * 1e-3) / 1e9
    tr = [traffic.get(k) for k in ("element_kernel", "gather_blocks_kernel", "gather_residual_kernel")]
    tr_sum = sum(tr) if all(v is not None for v in tr) else None
    roofline = {"bound": "hbm", "kernel": "assembly = element_kernel + gather_blocks_kernel + gather_residual_kernel",
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": tr_sum,
                "traffic_unit": "bytes per step (ncu dram read+write of the three kernels, profiles/ncu_traffic.json, quoted only while its source hash matches the kernels that ran)",
                # how busy the memory system is with the bytes the kernels REALLY move (K_e staging included)
                "dram_frac_on_measured_traffic": (tr_sum / (asm_ms * 1e-3) / 1e9 / hbm_peak) if tr_sum else None,
                "peak_source": f"MEASURED_PEAKS.json ({peaks_src})",
                "algorithmic_bytes_per_element": BYTES_PER_ELEM, "kernel_ms": asm_ms,
                "kernel_ms_samples": ph.get("phase_samples"),
                "phase_ms": {"element": elem_ms, "gather_k": gather_ms, "gather_r": gres_ms, "bc": bc_ms, "halo_x": halo_x_ms},
                "per_kernel_dram_traffic": {k: traffic.get(k) for k in ("element_kernel", "gather_blocks_kernel", "gather_residual_kernel")}}
    roofline_fp64 = {"bound": "fp64", "achieved": FLOP_PER_ELEM * cnt["local_elems"] / (asm_ms * 1e-3) / 1e12,
                     "peak": dfma_tf, "unit": "TFLOP/s", "peak_source": "DFMA probe measured in this run",
                     "flop_per_element": FLOP_PER_ELEM, "dmma_m8n8k4_tflops_this_run": dmma_tf,
                     "note": "DMMA (mma.sync.m8n8k4.f64) measured beside DFMA: the tensor path is used only if it beats the FMA pipe"}
    roofline_fp64["frac"] = roofline_fp64["achieved"] / dfma_tf if dfma_tf else None
    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "fea_gpu_step_from_host (pinned host nodes in, host residual out)"},
            "roofline": roofline, "roofline_fp64": roofline_fp64, "parity": parity, "comm": comm,
            "wall_s_timed_region": wall, "bad_points": bad, "residual_max_abs": r_max,
            "measured_copy_gbs_this_run": copy_gbs}
    if newton:
        line["newton"] = newton
        sp_ach = (76.0 * cnt["nnzb"] + 60.0 * cnt["owned_nodes"]) / (newton["spmv_ms"] * 1e-3) / 1e9
        line["roofline_spmv"] = {"bound": "hbm", "kernel": "spmv_sell_kernel", "achieved": sp_ach, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": sp_ach / hbm_peak, "traffic": traffic.get("spmv_sell_kernel"),
                                 "algorithmic_bytes": "76*nnzb + 20*n (BSR form of SURVEY 8d)",
                                 "kernel_ms": newton["spmv_ms"], "timed": "CUDA events around every in-solve SpMV launch"}
    if strong:
        line["strong_c4"] = strong

    if world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_baseline_sample
        v, sec, ne, kind = cpu_sample(n_s, args.model, 1, 0, 1)
        line["cpu_baseline"] = {"value": v, "unit": "elements/s", "cores": 1, "kind": kind,
                                "sample": f"one pass over a Kuhn sub-block of {n_s}^3 cubes ({ne} tets) in the same state, "
                                          f"{sec:.1f} s on one host core (the reference is single-threaded)"}
        if not args.no_newton:
            line["cpu_baseline"]["newton"] = cpu_newton_iteration(args.cpu_newton_sample, args.model)
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not parity["ok"]:
        log("[bench] PARITY CHECK FAILED")
        sys.exit(1)


def strong_c4(args, fg, rank, world, local_rank, new_nccl_id, barrier, allmax, allsum):
    """BASELINE configs[3]: one fixed Kuhn cube (default 75^3 -> 2 531 250 tets, 10 328 853 DOF), Neo-Hookean,
    partitioned over the N ranks by recursive coordinate bisection (2 x 2 x 2 boxes at N = 8).  One Newton
    iteration of the first load increment: assembly, Jacobi-PCG to the reference's tolerance, update."""
    n = args.c4_n
    t0 = time.time()
    mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 2, 0.005)
    g = fg.FeaGpu(mb["nodes"], mb["conn"], 1, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"],
                  rank=rank, nranks=world, nccl_id=new_nccl_id(), device=local_rank)
    cnt = g.counts()
    setup_s = allmax(time.time() - t0)
    g.apply_increment(1.0)
    g.update_nodes()                        # warm, the halo exchange included (NCCL sets its peer channels up on first use)
    g.assemble_all(True, fuse_bc=True)
    g.sync()
    g.phase_ms()
    barrier()
    g.timer_start()
    for _ in range(3):
        g.update_nodes()
        g.assemble_all(True, fuse_bc=True)
    asm_ms = allmax(g.timer_stop()) / 3
    ph = g.phase_ms()
    barrier()
    g.sync()
    g.timer_start()
    g.assemble_all(True)
    g.apply_bc(0.0)
    it, rr, ok = g.solve(args.lin_tol, args.lin_max_iter, fg.X0_ZERO, allow_unconverged=True)
    tol = g.dot_R_u()
    g.update_nodes()
    nt_ms = allmax(g.timer_stop())
    p = g.phase_ms()
    # the same system under the task files' PCG_ILU setting (Chebyshev-accelerated Jacobi)
    g.set_param("precond", 1)
    barrier()
    g.sync()
    g.timer_start()
    it_c, rr_c, ok_c = g.solve(args.lin_tol, args.lin_max_iter, fg.X0_ZERO, allow_unconverged=True)
    cheb_ms = allmax(g.timer_stop())
    pc = g.phase_ms()
    tol_c = g.dot_R_u()
    g.set_param("precond", 0)
    halo_ms, allreduce_ms = g.bench_comm(50)
    bad = g.bad_points()
    out = {"workload": f"Kuhn cube {n}^3, Neo-Hookean compressible, lambda=mu=100, bc: y faces prescribed in y + two pinned corners "
                       f"(no rigid mode), increment 0.005 h; BASELINE configs[3]",
           "elements": int(cnt["global_elems"]), "dof": 3 * int(cnt["global_nodes"]), "n_gpus": world,
           "partition": "recursive coordinate bisection", "neighbours_max": int(allmax(float(cnt["neighbours"]))),
           "local_elements_max": int(allmax(float(cnt["local_elems"]))), "halo_nodes_recv_max": int(allmax(float(cnt["halo_recv"]))),
           "assembly_ms": asm_ms, "element_assemblies_per_sec": cnt["global_elems"] / (asm_ms * 1e-3),
           "assembly_phase_ms": {"element": allmax(ph["element"]), "gather_k": allmax(ph["gather_k"]), "halo_x": allmax(ph["halo"])},
           "newton_iter_ms": nt_ms, "newton_iters_per_sec": 1e3 / nt_ms, "pcg_iters": it, "pcg_relres": rr,
           "pcg_exit": p["pcg_exit"], "pcg_ms_per_iter": allmax(p["pcg"]) / max(it, 1), "spmv_ms": allmax(p["spmv_avg"]),
           "halo_ms": allmax(halo_ms), "allreduce_ms": allmax(allreduce_ms), "dot_R_u": tol, "bad_points": bad,
           "pcg_ilu_setting": {"preconditioner": "Chebyshev-accelerated Jacobi, degree 4 on [lmax/100, lmax] of D^-1 A", "pcg_iters": it_c,
                               "pcg_relres": rr_c, "pcg_exit": pc["pcg_exit"], "solve_ms": cheb_ms, "jacobi_solve_ms": allmax(p["pcg"]),
                               "iteration_ratio": it / max(it_c, 1), "dot_R_u": tol_c},
           "setup_s": setup_s}
    if rank == 0:
        log(f"[bench] strong_c4: {json.dumps(out)}")
    g.close()
    return out


if __name__ == "__main__":
    main()
