"""GPU tuning helper (not part of the product path): SpMV lanes-per-row sweep and phase times."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import numpy as np
import fea_gpu as fg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, 1, 0.01)
g = fg.FeaGpu(mb["nodes"], mb["conn"], 0, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
cnt = g.counts()
g.apply_increment(1.0); g.assemble_all(True); g.apply_bc(0.0)
bytes_ = 76.0 * cnt["nnzb"] + 60.0 * cnt["owned_nodes"]
print("sell padding:", cnt["sell_slots"] / cnt["nnzb"])
for rep in range(3):
    ms = g.bench_spmv(50)
    print(f"spmv: {ms:.4f} ms  {bytes_ / ms / 1e6:.0f} GB/s (algorithmic BSR bytes)")
