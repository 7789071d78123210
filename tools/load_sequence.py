"""Full-Newton load sequence on a Kuhn block under torchrun (tools; results go to profiles/).

    python -m torch.distributed.run --nproc-per-node N ... tools/load_sequence.py <n> <increments> [model] [bc_style]
    env: SEQ_DY (increment, default 0.2 = 20 % of a cell), SEQ_LIN_TOL (PCG tolerance, default 1e-12),
         SEQ_TOTAL (total top-face displacement: <increments> is then ignored and the schedule is SEQ_DY0 (default
         0.2), doubling every increment up to SEQ_DY, until the total is reached -- needs SEQ_PREDICTOR=1, which
         scales its extrapolation by the ratio of consecutive increments), SEQ_PRECOND=1 (Chebyshev-Jacobi PCG),
         SEQ_PREDICTOR=1 (start every increment from x + (x - x_previous_increment): the boundary nodes move
         by the same increment every time, so this is the reference's boundary move plus a secant guess for the
         interior; the equilibrium Newton converges to is the same, it just starts closer),
         SEQ_CHECK_EVERY (sigma check / log line every k increments, default 1)

Unit-cube cells, "analytical" boundary set (the homogeneous uniaxial state is the exact solution,
exact-solutions/uniaxial) with the rigid rotation about y removed (bc_style 2; style 0 = exactly
the reference's set, whose singular K is numerically fragile beyond ~10 M DOF).  The reference moves
the boundary nodes by the full increment before equilibrating (fea_solver.c:168), so an increment
must stay below the cell size: 0.2 h here, the ratio of the shipped bricks (0.05 on 0.25 cells).
Prints per increment: Newton iterations, PCG iterations, sigma_yy error against the closed form."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import numpy as np
import fea_gpu as fg

n, increments = int(sys.argv[1]), int(sys.argv[2])
model = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bc_style = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
nccl_id, dist = None, None
if world > 1:
    import torch, torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [fg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0); nccl_id = box[0]


def closed_form(k1, lam=100.0, mu=100.0):
    if model == 0:
        kk2 = (3 * lam + 2 * mu - lam * k1 * k1) / (2 * lam + 2 * mu)
        return kk2 ** 0.5, (k1 / kk2) * ((lam + 2 * mu) * k1 * k1 + 2 * lam * kk2 - (3 * lam + 2 * mu)) / 2
    k2 = 1.0
    for _ in range(80):
        k2 -= (mu * (k2 * k2 - 1) + lam * np.log(k1 * k2 * k2)) / (2 * mu * k2 + 2 * lam / k2)
    J = k1 * k2 * k2
    return k2, (mu * (k1 * k1 - 1) + lam * np.log(J)) / J


L = float(n)
t0 = time.time()
DY = float(os.environ.get("SEQ_DY", "0.2"))   # per increment; default 20 % of a cell, the ratio of the shipped bricks (0.05 on 0.25-size cells)
LIN_TOL = float(os.environ.get("SEQ_LIN_TOL", "1e-12"))
PREDICTOR = os.environ.get("SEQ_PREDICTOR", "0") == "1"
CHECK_EVERY = int(os.environ.get("SEQ_CHECK_EVERY", "1"))
mb = fg.mesh_block(n, n, n, L, L, L, 0.0, bc_style, DY)
g = fg.FeaGpu(mb["nodes"], mb["conn"], model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"],
              rank=rank, nranks=world, nccl_id=nccl_id, device=int(os.environ.get("LOCAL_RANK", 0)))
cnt = g.counts()
if rank == 0:
    print(f"setup {time.time() - t0:.1f}s: {len(mb['conn'])} tets, {3 * len(mb['nodes'])} DOF, {world} rank(s), rank0 {cnt}", flush=True)
if os.environ.get("SEQ_PRECOND", "0") == "1":
    g.set_param("precond", 1)
TOTAL = float(os.environ.get("SEQ_TOTAL", "0"))
if TOTAL > 0:      # ramped schedule: the first (plain boundary move) increments stay small, later ones ride on the predictor
    sizes, s = [], float(os.environ.get("SEQ_DY0", "0.2"))
    while sum(sizes) < TOTAL - 1e-12:
        sizes.append(min(s, DY, TOTAL - sum(sizes)))
        s *= 2.0
    increments = len(sizes)
else:
    sizes = [DY] * increments
if rank == 0:
    print(f"schedule: {increments} increments, sizes {sizes[:6]} ... {sizes[-2:]}, total {sum(sizes):.4f} (stretch {1 + sum(sizes) / L:.4f})", flush=True)
log = []
done_disp = 0.0
sample = np.arange(0, len(mb["conn"]), 53, dtype=np.int32)    # elements whose sigma_yy is checked (every rank: the ones it holds)
t_all = time.time()


def anybad():
    b = float(g.bad_points())
    if dist is not None:
        t = torch.tensor([b], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); b = float(t[0])
    return b > 0


for step in range(1, increments + 1):
    g.sync(); ts = time.time()
    size = sizes[step - 1]
    if PREDICTOR and step > 1:
        # x <- x_k + (s_k / s_{k-1}) (x_k - x_{k-1}): the boundary nodes move by exactly this increment, saved <- x_k
        g.extrapolate_nodes(size / sizes[step - 2])
        g.update_state()
        if anybad():                             # fall back to the plain boundary move
            g.restore_nodes(); g.save_nodes(); g.apply_increment(size / DY)
    else:
        g.save_nodes()
        g.apply_increment(size / DY)
    done_disp += size
    its, pcg = 0, []
    while True:
        its += 1
        g.assemble_all(True); g.apply_bc(0.0)
        it, rr, ok = g.solve(LIN_TOL, 40000, fg.X0_ZERO, allow_unconverged=True)
        tol = g.dot_R_u(); g.update_nodes(); pcg.append(it)
        if rank == 0:
            p = g.phase_ms()
            print(f"  newton {its}: pcg {it} its relres {rr:.2e} exit {p['pcg_exit']} <R,u> {tol:.3e}", flush=True)
        if abs(tol) <= 1e-12 * L ** 3 or its >= 12 or not np.isfinite(tol):
            break
    g.update_state(); g.sync()
    dt = time.time() - ts
    if step % CHECK_EVERY and step != increments:
        if rank == 0:
            print(json.dumps(dict(step=step, newton_iters=its, pcg_iters=pcg, last_R_dot_u=tol, seconds=dt)), flush=True)
        continue
    k1 = 1.0 + done_disp / L
    k2, sig = closed_form(k1)
    F, S, found = g.get_state_elems(sample)    # every rank checks the sampled elements it holds
    err = float(np.abs(S[found][:, :, 1, 1] / sig - 1).max()) if found.any() else 0.0
    bad = g.bad_points()
    if dist is not None:
        t = torch.tensor([err, float(bad)], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); err, bad = float(t[0]), float(t[1])
    rec = dict(step=step, stretch=k1, newton_iters=its, pcg_iters=pcg, last_R_dot_u=tol, sigma_yy_closed_form=sig,
               sigma_yy_max_rel_err=err, seconds=dt, bad_points=bad)
    log.append(rec)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if bad > 0 or its >= 12 or not np.isfinite(tol):      # same on every rank (bad is all-reduced, tol is a global dot product)
        if rank == 0:
            print("ABORT: the sequence left the convergent path", flush=True)
        break
if rank == 0:
    print("SUMMARY", json.dumps(dict(n=n, tets=len(mb["conn"]), dof=3 * len(mb["nodes"]), ranks=world, model=model, dy=DY, lin_tol=LIN_TOL,
                                     predictor=PREDICTOR, total_seconds=time.time() - t_all, final_stretch=1.0 + done_disp / L, increments=increments, newton_total=sum(r['newton_iters'] for r in log) if CHECK_EVERY == 1 else None,
                                     worst_sigma_yy_rel_err=max(r["sigma_yy_max_rel_err"] for r in log),
                                     steps=log)), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
