"""Same global bar solved with WORLD_SIZE ranks (tools only): prints PCG iterations and exit state."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import numpy as np, fea_gpu as fg
n, ny = int(sys.argv[1]), int(sys.argv[2])
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
nccl_id = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [fg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0); nccl_id = box[0]
mb = fg.mesh_block(n, ny, n, float(n), float(ny), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], 0, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"],
              rank=rank, nranks=world, nccl_id=nccl_id, device=int(os.environ.get("LOCAL_RANK", 0)))
g.apply_increment(1.0); g.assemble_all(True); g.apply_bc(0.0)
rng = np.random.default_rng(1); v = rng.standard_normal(3 * len(mb["nodes"]))
y = g.spmv(v)
t = time.time(); it, rr, ok = g.solve(1e-14, 30000, fg.X0_ZERO, allow_unconverged=True); dt = time.time() - t
p = g.phase_ms(); u = g.get_solution()
if rank == 0:
    print(f"WORLD {world}: its {it} relres {rr:.3e} exit {p['pcg_exit']} best {p['pcg_best_relres']:.3e} stall {p['pcg_stall']} "
          f"|u| {np.linalg.norm(u):.12e} |Kv| {np.linalg.norm(y):.12e} v.Kv {v @ y:.12e} ({dt:.1f}s)", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
