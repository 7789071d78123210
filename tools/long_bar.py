"""PCG behaviour on slender bars (tools only)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import numpy as np, fea_gpu as fg
n, aspect = int(sys.argv[1]), int(sys.argv[2])
model = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mb = fg.mesh_block(n, n * aspect, n, float(n), float(n * aspect), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
g.apply_increment(1.0); g.assemble_all(True); g.apply_bc(0.0)
for max_iter in (500, 2000, 6000, 20000, 60000):
    t = time.time(); it, rr, ok = g.solve(1e-14, max_iter, fg.X0_ZERO, allow_unconverged=True); dt = time.time() - t
    p = g.phase_ms()
    print(f"max_iter {max_iter}: its {it} relres {rr:.3e} ok {ok} exit {p['pcg_exit']} best {p['pcg_best_relres']:.3e} last {p['pcg_last_relres']:.3e} stall {p['pcg_stall']} ({dt:.1f}s)", flush=True)
    if p["pcg_exit"] != 0: break
