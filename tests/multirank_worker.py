"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multirank.py or by hand):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multirank_worker.py
Every rank drives its fea_gpu context through a Newton step (NCCL halo exchange + all-reduced
dots); rank 0 repeats the computation on a single-rank context and on the CPU oracle.
bench.py runs the same check (newton_step_check) through its own process group before the timed loop,
so the driver's 2/4/8-GPU scaling runs carry this evidence too."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def small_bar(world):
    """4 x (4 world + 1) x 4 cubes of edge 0.25, clamped ends, Neo-Hookean: every rank gets a slab."""
    from oracle.kuhn import kuhn_block
    from oracle.oracle import Model
    ny = 4 * world + 1
    mb = kuhn_block(4, ny, 4, 1.0, ny / 4.0, 1.0, 0.0, 1, 0.02)
    return Model(nodes=mb["nodes"], conn=mb["conn"], presc_node=mb["presc_node"], presc_type=mb["presc_type"],
                 presc_vals=mb["presc_vals"], model=1, lam=100.0, mu=100.0, gauss=5)


def newton_step_check(dist, rank, world, local, log=print):
    """One Newton step on `world` ranks vs one rank vs the CPU oracle.  Collective; returns
    (ok, errors) on rank 0 and (True, {}) elsewhere."""
    import torch
    import fea_gpu as fg
    from oracle.oracle import PortOracle
    box = [fg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    m = small_bar(world)
    rng = np.random.default_rng(3)
    x0 = m.nodes + 0.004 * rng.standard_normal(m.nodes.shape)
    g = fg.FeaGpu(m.nodes, m.conn, m.model, m.lam, m.mu, 5, m.presc_node, m.presc_type, m.presc_vals,
                  rank=rank, nranks=world, nccl_id=box[0], device=local)
    cnt = g.counts()
    assert cnt["neighbours"] >= 1 and cnt["halo_recv"] > 0
    g.set_nodes(x0)
    g.apply_increment(1.0)
    g.assemble_all(True)
    R0 = g.get_forces()                      # collective all-gather
    g.apply_bc(0.0)
    it, rr, ok = g.solve(1e-13, 20000)       # classic PCG: p.Ap, then (r.z, r.r) all-reduced every iteration
    assert ok and rr <= 1e-13, (it, rr)      # the reported residual must be the converged one on every rank
    tol = g.dot_R_u()
    u = g.get_solution()
    # the single-reduction (Chronopoulos-Gear) recurrences over the same communicator, with the halo in stream
    # order and beside the interior slices.  They may break down where classic PCG converges (the solve then
    # ends on its guard: error code, not a wrong answer), so they are compared only when they converged.
    alt = {}
    for name, overlap in (("single_reduction", 0), ("single_reduction_overlap", 1)):
        g.set_param("pcg_variant", 1)
        g.set_param("pcg_overlap", overlap)
        it_a, rr_a, ok_a = g.solve(1e-13, 20000, allow_unconverged=True)
        alt[name] = (it_a, rr_a, ok_a, g.get_solution())
    g.set_param("pcg_variant", 0)
    g.set_param("pcg_overlap", 0)
    it, rr, ok = g.solve(1e-13, 20000)       # back to the default for the update below
    assert np.array_equal(g.get_solution(), u), "multi-rank PCG is not bit-reproducible"
    g.update_nodes()                         # x += u, then halo exchange of x
    g.assemble_all(True)
    x1 = g.get_nodes()
    R1 = g.get_forces()
    F, S = g.get_state()                     # only owned elements are written
    t = torch.from_numpy(S.copy()); dist.all_reduce(t); Ssum = t.numpy()   # each element owned once
    # host-buffer step on every rank (each fills its owned rows of the shared-shape host array)
    xh, Rh = fg.host_array(m.nodes.shape), fg.host_array(m.n_dof)
    xh[:] = x1; Rh[:] = 0.0
    g.step_from_host(xh, Rh, True)
    t = torch.from_numpy(np.array(Rh)); dist.all_reduce(t); R_host = t.numpy()
    good, e = True, {}
    if rank == 0:
        s1 = fg.FeaGpu(m.nodes, m.conn, m.model, m.lam, m.mu, 5, m.presc_node, m.presc_type, m.presc_vals, device=local)
        s1.set_nodes(x0); s1.apply_increment(1.0); s1.assemble_all(True)
        e = {"R0": relmax(R0, s1.get_forces())}
        s1.apply_bc(0.0)
        it1, rr1, ok1 = s1.solve(1e-13, 20000)
        e["u"] = relmax(u, s1.get_solution())
        for name, (it_a, rr_a, ok_a, u_a) in alt.items():
            e["u_" + name] = relmax(u_a, u) if ok_a else 0.0
        e["tol"] = abs(tol - s1.dot_R_u()) / abs(tol)
        s1.update_nodes(); s1.assemble_all(True)
        e["x1"] = relmax(x1 - m.nodes, s1.get_nodes() - m.nodes)
        e["R1"] = relmax(R1, s1.get_forces())
        e["S"] = relmax(Ssum, s1.get_state()[1])
        s1.apply_bc(0.0)
        e["R_hostpath"] = relmax(R_host, s1.get_forces())
        o = PortOracle(m)
        o.set_nodes(x0); o.apply_increment(1.0); o.update_state(); o.assemble_stiffness(); o.assemble_residual()
        e["R0_oracle"] = relmax(R0, o.get_forces())
        o.apply_bc(0.0); o.solve_slae()
        e["u_oracle"] = relmax(u, o.get_solution())
        alt_txt = ", ".join(f"{k} {v[0]} its{'' if v[2] else ' (ended on its guard)'}" for k, v in alt.items())
        log(f"MULTIRANK {world} ranks, pcg its {it} (one rank {it1}; {alt_txt}) errors "
            + str({k: f"{v:.2e}" for k, v in e.items()}))
        good = ok and ok1 and abs(it - it1) <= max(3, it1 // 50) \
            and e["R0"] < 1e-12 and e["R1"] < 1e-9 and e["u"] < 1e-9 and e["u_single_reduction"] < 1e-9 \
            and e["u_single_reduction_overlap"] < 1e-9 and e["x1"] < 1e-9 and e["S"] < 1e-9 and e["R0_oracle"] < 1e-12 \
            and e["u_oracle"] < 1e-9 and e["tol"] < 1e-9 and e["R_hostpath"] < 1e-9
        e["pcg_iters"] = {"ranks": it, "one_rank": it1, **{k: (v[0] if v[2] else -v[0]) for k, v in alt.items()}}
        s1.close()
    g.close()
    return bool(good), e


def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    good, _ = newton_step_check(dist, rank, world, local, log=lambda s: print(s, flush=True))
    if rank == 0:
        print("MULTIRANK_RESULT", "PASS" if good else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
