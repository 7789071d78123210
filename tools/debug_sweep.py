import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, fea_gpu as fg
from conftest import block_model
from oracle.oracle import PortOracle, uniaxial_neohookean, uniaxial_a5
model = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = block_model((3, 3, 3), model=model, bc_style=0, dy=1.0 / 120)
g = fg.FeaGpu(m.nodes, m.conn, m.model, m.lam, m.mu, 5, m.presc_node, m.presc_type, m.presc_vals)
o = PortOracle(m)
for step in range(1, 4):
    g.apply_increment(1.0); o.apply_increment(1.0)
    for it in range(1, 9):
        g.assemble_all(True); g.apply_bc(0.0)
        o.update_state(); o.assemble_stiffness(); o.assemble_residual(); o.apply_bc(0.0)
        Rg, Ro = g.get_forces(), o.get_forces()
        its, rr, ok = g.solve(1e-14, 20000)
        ito = o.solve_slae(1e-15, 200000)
        ug, uo = g.get_solution(), o.get_solution()
        tol_g, tol_o = g.dot_R_u(), o.dot_forces_solution()
        res_g = np.linalg.norm(g.spmv(ug) - Rg) / max(np.linalg.norm(Rg), 1e-300)
        print(f"step {step} it {it}: |R| g {np.linalg.norm(Rg):.3e} o {np.linalg.norm(Ro):.3e} | pcg its g {its} (rr {rr:.1e} ok {ok}) o {ito} | "
              f"|u| g {np.linalg.norm(ug):.3e} o {np.linalg.norm(uo):.3e} | true relres g {res_g:.1e} | <R,u> g {tol_g:.2e} o {tol_o:.2e}")
        g.update_nodes(); o.update_with_solution()
        if abs(tol_g) < 1e-20 and abs(tol_o) < 1e-20: break
    g.update_state(); o.update_state()
    k1 = 1 + step / 120; k2, sig = (uniaxial_neohookean if model else uniaxial_a5)(k1)
    Fg, Sg = g.get_state(); Fo, So = o.get_state()
    print(f"  == step {step}: syy relerr g {np.abs(Sg[:,:,1,1]/sig-1).max():.1e} o {np.abs(So[:,:,1,1]/sig-1).max():.1e}; F00 relerr g {np.abs(Fg[:,:,0,0]/k2-1).max():.1e} o {np.abs(Fo[:,:,0,0]/k2-1).max():.1e}")
