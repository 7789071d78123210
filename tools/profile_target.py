"""Small fixed sequence of the hot kernels for ncu captures (tools only, not product path)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import fea_gpu as fg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], 0, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
g.apply_increment(1.0)
for _ in range(2):
    g.assemble_all(True)
    g.apply_bc(0.0)
for _ in range(2):
    g.assemble_all(True, fuse_bc=True)     # what a bench step runs
g.bench_spmv(2)
g.solve(1e-14, 4, fg.X0_ZERO, allow_unconverged=True)
g.sync()
print("profile target done", g.counts())
