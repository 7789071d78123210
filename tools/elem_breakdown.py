import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import fea_gpu as fg
n = 55
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
for model in (0, 1):
    g = fg.FeaGpu(mb["nodes"], mb["conn"], model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
    g.apply_increment(1.0)
    for occ in ("",):
        for name, fn in (("state only (K=0,R=0)", g.update_state), ("residual (K=0,R=1)", g.assemble_residual),
                         ("stiffness (K=1,R=0)", g.assemble_stiffness), ("all (K=1,R=1)", lambda: g.assemble_all(True))):
            fn(); fn(); fn()
            print(f"model {model} {name:24s} element kernel {g.phase_ms()['element']:.3f} ms", flush=True)
    g.close()
