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
steps, and reported under "newton".

Workload at N GPUs: Kuhn block of n x (n*N) x n cubes of unit size (weak scaling, n = 55
-> 998 250 ten-node tets and 4 102 893 DOF per GPU = BASELINE.json configs[2]), A5
compressible, lambda = mu = 100, 5-point rule, deformed by the exact uniaxial map at
stretch 1.5 plus a seeded perturbation of 1e-3 h.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))

FLOP_PER_ELEM = 29000.0     # index form, 5-point rule, FMA = 2
BYTES_PER_ELEM = 3600.0     # compulsory HBM traffic of a fused assembly
METRIC = "element_assemblies_per_sec"


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
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


def deformed_state(nodes, h, seed=12345, k1=1.5, model=0):
    """Exact uniaxial map at stretch k1 plus a uniform perturbation of amplitude 1e-3 h."""
    k2 = lateral_stretch(k1, model)
    rng = np.random.default_rng(seed)
    x = nodes * np.array([k2, k1, k2])
    x += (rng.random(nodes.shape) - 0.5) * 2e-3 * h
    return x


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own compiled element code (oracle/_ref) or
# the plain-C port, on a bounded sample of the same workload


def _cpu_sample_worker(args):
    n_s, model, steps, warmup, seed = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
    import fea_gpu as fg
    from oracle.oracle import Model, PortOracle, RefOracle, have_ref
    mb = fg.mesh_block(n_s, n_s, n_s, float(n_s), float(n_s), float(n_s), 0.0, 0, 0.0)
    m = Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=mb["presc_node"], presc_type=mb["presc_type"],
              presc_vals=mb["presc_vals"], model=model, lam=100.0, mu=100.0, gauss=5)
    kind = "reference" if have_ref() else "port"
    o = RefOracle(m) if kind == "reference" else PortOracle(m)
    o.set_nodes(deformed_state(m.nodes, 0.5, seed, model=model))
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
    sample = (f"{cores} independent Kuhn sub-blocks of {n_s}^3 cubes ({ne} tets each), same state and pass as the "
              f"GPU arm; reference objects compiled from /root/reference with a sorted-array stand-in for libspmatrix"
              if kind == "reference" else f"{cores} x {ne} tets, plain-C port of the reference (oracle/oracle_fea.c)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus, note="bounded sample"),
            "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, world, note=None):
    n = args.n
    cfg = {"workload": f"Kuhn block {n}x{n * world}x{n} cubes, 10-node tets, "
                       f"{'A5' if args.model == 0 else 'Neo-Hookean'} compressible, lambda=mu=100, 5-pt rule "
                       f"(BASELINE configs[2] per GPU)",
           "elements": 6 * n * n * n * world, "dof": 3 * (2 * n + 1) * (2 * n * world + 1) * (2 * n + 1),
           "parallelism": f"row/element partition over {world} GPU(s), slabs along y",
           "l2": "working set (K_e staging + matrix values, >7 GB per GPU) exceeds L2; no flush needed",
           "state": "uniaxial map at stretch 1.5 + 1e-3 h perturbation, seed 12345",
           "bc": "faces y=min / y=max clamped (type 7, as data/*_brick.sexp), 0.01 h per increment"}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--n", type=int, default=55, help="cubes per edge per GPU (55 -> 998 250 tets)")
    ap.add_argument("--model", type=int, default=0, help="0 = A5 (configs[2]), 1 = Neo-Hookean")
    ap.add_argument("--newton-iters", type=int, default=2)
    ap.add_argument("--lin-tol", type=float, default=1e-14, help="PCG relative tolerance (reference files: 1e-14)")
    ap.add_argument("--lin-max-iter", type=int, default=20000)
    ap.add_argument("--ref-sample", type=int, default=10, help="cubes per edge of the CPU sample")
    ap.add_argument("--cpu-baseline-sample", type=int, default=14)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import fea_gpu as fg

    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist   # host-side plumbing only (gloo): id broadcast, barrier, max
        dist.init_process_group("gloo", rank=rank, world_size=world)
        box = [fg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    n = args.n
    t0 = time.time()
    mb = fg.mesh_block(n, n * world, n, float(n), float(n * world), float(n), 0.0, 1, 0.01)
    nodes, conn = mb["nodes"], mb["conn"]
    n_elems, n_dof = len(conn), 3 * len(nodes)
    x_def = deformed_state(nodes, 0.5, model=args.model)
    g = fg.FeaGpu(nodes, conn, args.model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"],
                  rank=rank, nranks=world, nccl_id=nccl_id, device=local_rank)
    cnt = g.counts()
    if rank == 0:
        log(f"[bench] mesh+plan+upload {time.time() - t0:.1f}s: {n_elems} tets, {n_dof} DOF, rank0 {cnt}")
    g.set_nodes(x_def)

    def step():
        g.update_nodes()            # x += u (u = 0 here) and, for N > 1, the halo exchange of x
        g.assemble_all(True, fuse_bc=True)   # geometry, F, stress, tangent, K_e, R_e, both gathers with the
                                             # Dirichlet cancellation of solver_apply_prescribed_bc(0) folded in

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
    elem_ms = gather_ms = 0.0
    for _ in range(args.steps):
        step()
    ms = g.timer_stop()             # CUDA events on the launching stream, synchronises
    barrier()
    wall = time.perf_counter() - wall0
    launches = fg.launch_count() - launches0
    ph = g.phase_ms()
    ms = allmax(ms)
    clocks = sampler.stop() if rank == 0 else None
    value = n_elems * args.steps / (ms * 1e-3)
    bad = g.bad_points()

    # per-kernel durations: CUDA events around every launch of the timed steps, averaged (no sync inside the loop)
    elem_ms, gather_ms, gres_ms, bc_ms = ph["element"], ph["gather_k"], ph["gather_r"], ph["bc"]

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
    e2e_value = n_elems * args.steps / e2e_s
    h2d, d2h = allsum(float(h2d)), allsum(float(d2h))
    r_check = float(np.abs(Rh).max())

    # ---- Newton iterations: assembly + PCG to the reference's tolerance + update -------------
    newton = None
    if not args.no_newton:
        g.set_nodes(nodes)
        its, nt_ms, relres, spmv_ms, exits = [], [], [], [], []
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
            if rank == 0:
                log(f"[bench] newton it {k}: {t_ms:.1f} ms, pcg {it} its (relres {rr:.2e}, ok={ok}), <R,u>={tol:.3e}, "
                    f"spmv {p['spmv_avg']:.3f} ms")
        nnzb = allsum(float(cnt["nnzb"]))
        spmv_bytes_rank = 76.0 * cnt["nnzb"] + 20.0 * 3 * cnt["owned_nodes"]   # BSR: SURVEY 8d
        sp = float(np.mean(spmv_ms))
        newton = {"newton_iters_per_sec": len(nt_ms) / (sum(nt_ms) * 1e-3), "ms_per_newton_iter": float(np.mean(nt_ms)),
                  "pcg_iters_per_newton_iter": float(np.mean(its)), "pcg_relres": float(max(relres)),
                  "pcg_tol": args.lin_tol, "pcg_exit": exits, "pcg_exit_legend": "1 = tolerance met, 2 = stall/divergence guard, 0 = max_iter",
                  "pcg_iters_per_sec": float(sum(its) / (sum(nt_ms) * 1e-3)),
                  "spmv_ms": sp, "spmv_format": "3x3 blocks in SELL-32-sigma, fp64 values, int32 block columns",
                  "nnz_scalar_total": 9 * nnzb}
    peaks, peaks_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dfma_tf, copy_gbs = fg.measure_peaks(local_rank)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # DRAM traffic per launch from the committed ncu --set full capture (same workload only)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t["workload"]["n"] == args.n and t["workload"]["n_gpus"] == world and args.model == 0:
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass

    # roofline of the dominant kernel of the timed step
    dom = "element_kernel" if elem_ms >= gather_ms else "gather_blocks_kernel"
    dom_ms = max(elem_ms, gather_ms)
    ach = BYTES_PER_ELEM * cnt["local_elems"] / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic.get(dom), "traffic_unit": "bytes per launch (ncu dram read+write)",
                "peak_source": f"MEASURED_PEAKS.json ({peaks_src})",
                "algorithmic_bytes_per_element": BYTES_PER_ELEM, "kernel_ms": dom_ms,
                "kernel_ms_samples": ph.get("phase_samples"),
                "phase_ms": {"element": elem_ms, "gather_k": gather_ms, "gather_r": gres_ms, "bc": bc_ms}}
    asm_ms = elem_ms + gather_ms + gres_ms
    roofline_fp64 = {"bound": "fp64", "achieved": FLOP_PER_ELEM * cnt["local_elems"] / (asm_ms * 1e-3) / 1e12,
                     "peak": dfma_tf, "unit": "TFLOP/s", "peak_source": "DFMA probe measured in this run",
                     "flop_per_element": FLOP_PER_ELEM}
    roofline_fp64["frac"] = roofline_fp64["achieved"] / dfma_tf if dfma_tf else None
    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "fea_gpu_step_from_host (pinned host nodes in, host residual out)"},
            "roofline": roofline, "roofline_fp64": roofline_fp64,
            "wall_s_timed_region": wall, "bad_points": bad, "residual_check_max": r_check,
            "measured_copy_gbs_this_run": copy_gbs}
    if newton:
        line["newton"] = newton
        sp_ach = (76.0 * cnt["nnzb"] + 60.0 * cnt["owned_nodes"]) / (newton["spmv_ms"] * 1e-3) / 1e9
        line["roofline_spmv"] = {"bound": "hbm", "kernel": "spmv_sell_kernel", "achieved": sp_ach, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": sp_ach / hbm_peak, "traffic": traffic.get("spmv_sell_kernel"),
                                 "algorithmic_bytes": "76*nnzb + 20*n (BSR form of SURVEY 8d)",
                                 "kernel_ms": newton["spmv_ms"], "timed": "CUDA events around every in-solve SpMV launch"}

    if world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_baseline_sample
        v, sec, ne, kind = cpu_sample(n_s, args.model, 1, 0, 1)
        line["cpu_baseline"] = {"value": v, "unit": "elements/s", "cores": 1, "kind": kind,
                                "sample": f"one pass over a Kuhn sub-block of {n_s}^3 cubes ({ne} tets) in the same state, "
                                          f"{sec:.1f} s on one host core (the reference is single-threaded)"}
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
