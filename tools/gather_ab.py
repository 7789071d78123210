"""ncu / timing target: one assembly pass per gather mapping on the C3 block (tools only, not product path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import fea_gpu as fg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
modes = [int(a) for a in sys.argv[2:]] or [1, 9]
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], 0, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
g.apply_increment(1.0)
for mode in modes:
    g.set_param("gather_mode", mode)
    for _ in range(3):
        g.assemble_all(True, fuse_bc=True)
    g.sync(); g.phase_ms()
    for _ in range(5):
        g.assemble_all(True, fuse_bc=True)
    p = g.phase_ms()
    print(f"mode {mode}: element {p['element']:.3f} ms, gather_k {p['gather_k']:.3f} ms", flush=True)
