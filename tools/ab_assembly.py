"""A/B timing of two builds of the library (FEA_GPU_LIB): element / gather phase times on C3 (tools only)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import fea_gpu as fg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
for model in (0, 1):
    g = fg.FeaGpu(mb["nodes"], mb["conn"], model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
    g.apply_increment(1.0)
    for _ in range(4):
        g.assemble_all(True, fuse_bc=True)
    g.phase_ms()
    g.timer_start()
    for _ in range(10):
        g.assemble_all(True, fuse_bc=True)
    ms = g.timer_stop() / 10
    p = g.phase_ms()
    print(f"{os.path.basename(fg.LIB_PATH)} model {model}: step {ms:.3f} ms  element {p['element']:.3f}  gather_k {p['gather_k']:.3f}  gather_r {p['gather_r']:.3f}", flush=True)
    g.close()
