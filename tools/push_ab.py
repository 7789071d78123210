"""A/B of the pull gathers (gather_mode 1) against the direct (push) assembly (gather_mode 2), whole and chunked
(tools only, not product path).

    python tools/push_ab.py check          bitwise comparison of K on a 12^3 block, all code paths
    python tools/push_ab.py time [n] [chunk_tiles ...]   step times on the C3 block (n = 55)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
import numpy as np
import fea_gpu as fg

what = sys.argv[1] if len(sys.argv) > 1 else "check"


def ctx(n, model=0, bc=1):
    mb = fg.mesh_block(n, n, n, float(n), float(n), float(n), 0.0, bc, 0.01)
    g = fg.FeaGpu(mb["nodes"], mb["conn"], model, 100.0, 100.0, 5, mb["presc_node"], mb["presc_type"], mb["presc_vals"])
    rng = np.random.default_rng(1)
    x = mb["nodes"] * np.array([1.0, 1.3, 0.9]) + 1e-2 * rng.standard_normal(mb["nodes"].shape)
    g.set_nodes(x)
    return g, mb


if what == "check":
    ok = True
    for model in (0, 1):
        g, mb = ctx(int(sys.argv[2]) if len(sys.argv) > 2 else 12, model)
        res = {}
        g.set_param("gather_sym", 0)
        g.assemble_all(True); ref0 = g.get_csr()[3].copy()
        g.assemble_all(True, fuse_bc=True); ref1 = g.get_csr()[3].copy()
        g.set_param("gather_sym", 1)
        g.assemble_all(True); s0 = np.array_equal(ref0, g.get_csr()[3])
        g.assemble_all(True, fuse_bc=True); s1 = np.array_equal(ref1, g.get_csr()[3])
        print(f"model {model}: pull, upper triangle + mirror stores == every slot its own list: {s0} (plain) {s1} (Dirichlet)", flush=True)
        ok &= s0 and s1
        for ch, ov in ((37, 1), (64, 0), (5, 1)):
            g.set_param("chunk_tiles", ch); g.set_param("chunk_overlap", ov)
            g.assemble_all(True); s0 = np.array_equal(ref0, g.get_csr()[3])
            g.assemble_all(True, fuse_bc=True); s1 = np.array_equal(ref1, g.get_csr()[3])
            g.assemble_stiffness(); s2 = np.array_equal(ref0, g.get_csr()[3])
            print(f"model {model}: pull in chunks of {ch} tiles (overlap {ov}) == unchunked: {s0} {s1} {s2}", flush=True)
            ok &= s0 and s1 and s2
        g.set_param("chunk_tiles", 0); g.set_param("chunk_overlap", 1)
        for tag, mode, chunk, fuse in (("pull", 1, 0, False), ("push", 2, 0, False), ("push chunk 37", 2, 37, False),
                                       ("pull bc", 1, 0, True), ("push bc", 2, 0, True), ("push bc chunk 64", 2, 64, True),
                                       ("push bc chunk 1", 2, 1, True)):
            g.set_param("gather_mode", mode)
            g.set_param("chunk_tiles", chunk)
            g.assemble_all(True, fuse_bc=fuse)
            res[tag] = (g.get_csr()[3].copy(), g.get_forces().copy())
        for a, b in (("pull", "push"), ("pull", "push chunk 37"), ("pull bc", "push bc"), ("pull bc", "push bc chunk 64"),
                     ("pull bc", "push bc chunk 1")):
            same = np.array_equal(res[a][0], res[b][0]) and np.array_equal(res[a][1], res[b][1])
            print(f"model {model}: {a:8s} == {b:18s}: {same}   max|K| {np.abs(res[a][0]).max():.3e}  nnz {res[a][0].size}", flush=True)
            ok &= same
        g.set_param("gather_mode", 1); g.set_param("chunk_tiles", 0); g.assemble_all(True)
        k1 = g.element_matrix(77)
        g.set_param("gather_mode", 2); g.assemble_all(True)
        k2 = g.element_matrix(77)
        print("element matrix readback equal:", np.array_equal(k1, k2), flush=True)
        ok &= np.array_equal(k1, k2)
        # stiffness only / solve through the new path
        g.set_param("chunk_tiles", 16)
        g.assemble_stiffness(); v1 = g.get_csr()[3].copy()
        g.set_param("gather_mode", 1); g.assemble_stiffness()
        ok &= np.array_equal(v1, g.get_csr()[3])
        print("assemble_stiffness chunked equal:", np.array_equal(v1, g.get_csr()[3]), flush=True)
        g.close()
    print("ALL EQUAL" if ok else "MISMATCH", flush=True)
    sys.exit(0 if ok else 1)

n = int(sys.argv[2]) if len(sys.argv) > 2 else 55
chunks = [int(a) for a in sys.argv[3:]] or [0, 296, 592, 1184]
g, mb = ctx(n)
xs = fg.host_array((len(mb["nodes"]), 3)); xs[:] = g.get_nodes()
Rs = fg.host_array((len(mb["nodes"]), 3))


def run(tag, reps=10):
    for _ in range(3):
        g.assemble_all(True, fuse_bc=True)
    g.sync(); g.phase_ms()
    g.timer_start()
    for _ in range(reps):
        g.assemble_all(True, fuse_bc=True)
    ms = g.timer_stop() / reps
    p = g.phase_ms()
    g.sync(); t0 = time.perf_counter()
    for _ in range(reps):
        g.step_from_host(xs, Rs, True)
    e2e = (time.perf_counter() - t0) / reps * 1e3
    print(f"{tag:28s} step {ms:.3f} ms ({len(mb['conn']) / ms / 1e3:.1f} M el/s)  element {p['element']:.3f}  gather_k {p['gather_k']:.3f}  "
          f"gather_r {p['gather_r']:.3f}  host-buffer step {e2e:.3f} ms", flush=True)


g.set_param("gather_mode", 1)
g.set_param("gather_sym", 0); run("pull, every slot its own list")
g.set_param("gather_sym", 1); run("pull, upper + mirror (default)")
for sp in [int(a) for a in os.environ.get("PULL_SPLITS", "").split(",") if a]:
    g.set_param("gather_split", sp); run(f"pull, upper + mirror, split {sp}")
g.set_param("gather_split", 8)
for spec in [a for a in os.environ.get("PULL_CHUNKS", "").split(",") if a]:     # tiles:overlap[:split]
    f = [int(v) for v in spec.split(":")]
    g.set_param("chunk_tiles", f[0]); g.set_param("chunk_overlap", f[1]); g.set_param("gather_split", f[2] if len(f) > 2 else 8)
    run(f"pull, chunks of {f[0]} tiles, overlap {f[1]}, split {f[2] if len(f) > 2 else 8}")
g.set_param("chunk_tiles", 0); g.set_param("gather_split", 8); g.set_param("chunk_overlap", 1)
if os.environ.get("PUSH_SKIP"):
    sys.exit(0)
sp_ms = g.bench_spmv(20) if hasattr(g, "bench_spmv") else float("nan")
print(f"spmv {sp_ms:.4f} ms", flush=True)
g.set_param("gather_mode", 2)
for dbg in [int(a) for a in os.environ.get("PUSH_DBG", "").split(",") if a]:
    g.set_param("cells_dbg", dbg)
    run(f"push, cells_dbg {dbg} (diagnostic)")
g.set_param("cells_dbg", 0)
for ch in chunks:
    g.set_param("chunk_tiles", ch)
    run(f"push, chunk_tiles {ch}")
