"""Generate the committed golden fixtures from the REFERENCE's own compiled code.

Run only where /root/reference exists (this container):

    make -C oracle ref && python tests/golden/make_golden.py

Everything written here comes out of oracle/_ref/libfea_ref.so, i.e. the reference's
fea_solver.c / fea_model.c / dense_matrix.c compiled from /root/reference (see
oracle/Makefile); the meshes are the reference's shipped model files
solver-large/data/*.sexp re-encoded as arrays (the .sexp files do not travel to the
GPU box).  Fixtures are kept small: the full K is pinned through seeded probe products
K.v plus a few dense element matrices; the entry-wise comparison happens live against
oracle/oracle_fea.c, which tests/test_oracle_golden.py pins against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle, load_sexp  # noqa: E402

DATA = "/root/reference/solver-large/data"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = ["neohook_brick", "a5_brick", "neohook_brick_analytical", "a5_brick_analytical"]
PROBE_ELEMS = [0, 5, 77, 200, 345]


def csr_mv(rp, ci, v, x):
    y = np.zeros(len(rp) - 1)
    np.add.at(y, np.repeat(np.arange(len(rp) - 1), np.diff(rp)), v * x[ci])
    return y


def main():
    rng = np.random.default_rng(20261018)
    for name in CASES:
        m = load_sexp(os.path.join(DATA, name + ".sexp"))
        probes = rng.standard_normal((4, m.n_dof))
        r = RefOracle(m)
        r.apply_increment(1.0)
        r.update_state()
        F, S = r.get_state()
        g, detJ = r.get_gradients()
        r.assemble_stiffness()
        r.assemble_residual()
        rp, ci, v = r.get_csr()
        R = r.get_forces()
        Kv = np.stack([csr_mv(rp, ci, v, p) for p in probes])
        diag = np.array([v[rp[i]:rp[i + 1]][ci[rp[i]:rp[i + 1]] == i][0] for i in range(m.n_dof)])
        ke = np.stack([r.element_matrix(e) for e in PROBE_ELEMS])
        r.apply_bc(0.0)
        rp2, ci2, v2 = r.get_csr()
        R_bc = r.get_forces()
        Kv_bc = np.stack([csr_mv(rp2, ci2, v2, p) for p in probes])
        r.solve_slae()
        u = r.get_solution()
        r.close()
        # the reference's solve() end to end, first two load increments
        rhs, sol, tol = RefOracle.run_solve(m, load_increments=2)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            nodes=m.nodes, conn=m.conn, presc_node=m.presc_node, presc_type=m.presc_type,
            presc_vals=m.presc_vals, model=m.model, lam=m.lam, mu=m.mu, gauss=m.gauss,
            desired_tolerance=m.desired_tolerance, modified_newton=int(m.modified_newton),
            max_newton=m.max_newton, solver_type=m.solver_type,
            F=F, S=S, detJ=detJ, g_sample=g[PROBE_ELEMS], R=R, rowptr=rp, nnz=len(v),
            colidx_head=ci[:2000], probes=probes, Kv=Kv, Kdiag=diag, probe_elems=PROBE_ELEMS, Ke=ke,
            R_bc=R_bc, Kv_bc=Kv_bc, u_first=u,
            newton_tol=tol, newton_u_head=sol[:3], newton_u_sum=sol.sum(axis=0), newton_count=len(sol))
        print(name, "nnz", len(v), "newton solves", len(sol), "final tol", tol[-1])
        if name == "neohook_brick_analytical":
            # the reference's own Gmsh export of those two increments (fea_solver.c:1375-1488)
            import gzip
            with open("/tmp/fea_ref_out.msh", "rb") as f, open(os.path.join(OUT, name + "_2steps.msh.gz"), "wb") as raw:
                with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as g:   # mtime=0: reproducible bytes
                    g.write(f.read())

    # fea_model.c on bare deformation gradients (both models)
    Fs = np.eye(3)[None] + 0.2 * rng.standard_normal((16, 3, 3))
    Fs = Fs[np.linalg.det(Fs) > 0.2]
    out = {"F": Fs}
    for model, tag in ((0, "a5"), (1, "nh")):
        S, Ct = zip(*[RefOracle.model_eval(model, 100.0, 100.0, F) for F in Fs])
        out["S_" + tag], out["C_" + tag] = np.stack(S), np.stack(Ct)
    for ng in (4, 5):
        gt, N, dN = RefOracle.tables(ng)
        out[f"gauss{ng}"], out[f"N{ng}"], out[f"dN{ng}"] = gt, N, dN
    np.savez_compressed(os.path.join(OUT, "model_tables.npz"), **out)
    print("model_tables:", len(Fs), "deformation gradients")




def brick_fine_fixture():
    """solver-large/data/brick_fine.sexp as arrays (34 070 nodes, 22 934 tets, unstructured), with
    the BC node ids shifted by -1: as shipped they are 1-based while the loader reads them 0-based
    (SURVEY 8c "fixture defect").  Reference-compiled probes pin K and R on the 1 % stretched state."""
    m = load_sexp(os.path.join(DATA, "brick_fine.sexp"))
    m.presc_node = (m.presc_node - 1).astype(np.int32)
    x = m.nodes.copy()
    x[:, 1] = 1.0 + (x[:, 1] - 1.0) * 1.01
    r = RefOracle(m)
    r.set_nodes(x)
    r.update_state()
    r.assemble_stiffness()
    r.assemble_residual()
    rp, ci, v = r.get_csr()
    rng = np.random.default_rng(77)
    probes = rng.standard_normal((2, m.n_dof))
    Kv = np.stack([csr_mv(rp, ci, v, p) for p in probes])
    F, S = r.get_state()
    # kept small: coordinates are 9-decimal in the file -> exact as integer nano-units; K and R are
    # pinned through every 40th entry of K.v / R plus their norms (the probe is regenerated from
    # its seed); the entry-wise check is done live against oracle/oracle_fea.c
    micro = np.rint(m.nodes * 1e9).astype(np.int64)       # the file carries 9 decimals
    assert np.array_equal(micro / 1e9, m.nodes)
    np.savez_compressed(os.path.join(OUT, "brick_fine.npz"), nodes_nano=micro,
                        conn=m.conn, presc_node=m.presc_node, presc_type=m.presc_type, presc_vals=m.presc_vals,
                        model=m.model, lam=m.lam, mu=m.mu, gauss=m.gauss, nnz=len(v), max_row=int(np.diff(rp).max()),
                        probe_seed=77, Kv_sample=Kv[:, ::40], Kv_norm=np.linalg.norm(Kv, axis=1),
                        R_sample=r.get_forces()[::40], R_norm=np.linalg.norm(r.get_forces()),
                        S_mean=S.mean(axis=(0, 1)), S_sample=S[::997])
    print("brick_fine nnz", len(v), "max row", np.diff(rp).max())


if __name__ == "__main__":
    main()
    brick_fine_fixture()
