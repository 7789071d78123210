"""GPU tuning helper (not part of the product path): element / gather phase times on C3 for the
gather scheduling knobs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import fea_gpu as fg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], 0, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
g.apply_increment(1.0)


def show(tag):
    for _ in range(3):
        g.assemble_all(True)
    p = g.phase_ms()
    print(f"{tag:40s} element {p['element']:.3f}  gather_k {p['gather_k']:.3f}  gather_r {p['gather_r']:.3f}", flush=True)


for gt in (128, 256):
    g.set_param("gather_threads", gt)
    for sp in (1, 2, 4, 6, 8):
        g.set_param("gather_split", sp)
        show(f"gather threads {gt} split {sp}")
