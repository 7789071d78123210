"""CPU, world_size 2 over gloo: the N>1 host logic of the library (fea_plan_*): node-range
partition, halo send/recv lists, owned-row matrices.  Each rank builds ITS plan only, fills
ghosts through the halo lists with real point-to-point messages, runs a distributed SpMV
and a Jacobi-PCG whose dot products are all-reduced -- the same communication pattern
fea_gpu_solve drives over NCCL -- and the result is compared with the single-rank oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, block_model, csr_mv
from oracle.oracle import PortOracle
from test_plan import bsr_to_dense_rows, gather_numpy, staged_blocks

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_exchange(dist, torch, p, vec):
    """vec: [n_local][3]; owners push interface entries, ghosts of one owner are contiguous."""
    reqs = []
    recv_bufs = []
    for i, q in enumerate(p.nbr_rank):
        send = torch.from_numpy(np.ascontiguousarray(vec[p.send_nodes[p.send_ptr[i]:p.send_ptr[i + 1]]]))
        reqs.append(dist.isend(send, int(q)))
        buf = torch.empty((int(p.recv_ptr[i + 1] - p.recv_ptr[i]), 3), dtype=torch.float64)
        reqs.append(dist.irecv(buf, int(q)))
        recv_bufs.append((i, buf))
    for r in reqs:
        r.wait()
    for i, buf in recv_bufs:
        vec[p.n_own + p.recv_ptr[i]: p.n_own + p.recv_ptr[i + 1]] = buf.numpy()


def _worker(rank, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fea-large_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import fea_gpu as fg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)

    m = block_model((2, 6, 2), model=1, bc_style=1, dy=0.02)
    p = fg.Plan(m.nodes, m.conn, rank, WORLD)
    rng = np.random.default_rng(5)
    x_glob = m.nodes + 0.01 * rng.standard_normal(m.nodes.shape)

    # 1. halo exchange of coordinates: ghosts must equal the owners' values
    x_loc = np.zeros((p.n_local, 3))
    x_loc[:p.n_own] = x_glob[p.node_gid[:p.n_own]]
    _halo_exchange(dist, torch, p, x_loc)
    assert np.array_equal(x_loc, x_glob[p.node_gid])

    # 2. owned rows of K from the rank's own elements (oracle arithmetic, numpy gather)
    o = PortOracle(type(m)(**{**m.__dict__, "conn": np.ascontiguousarray(m.conn[p.elem_gid])}))
    o.set_nodes(x_glob); o.update_state()
    vals = gather_numpy(p, staged_blocks(o, p.n_elems))
    A = bsr_to_dense_rows(p, vals, m.n_dof)            # [3 n_own][n_dof], global columns
    rows = (3 * p.node_gid[:p.n_own, None] + np.arange(3)).ravel()
    cols_local = (3 * p.node_gid[:, None] + np.arange(3)).ravel()
    A_loc = A[:, cols_local]                           # columns in local numbering (owned + ghosts)
    # make it SPD-solvable: clamp the two end faces (rows/cols of prescribed DOFs -> identity)
    presc = np.zeros(m.n_dof, bool)
    presc[(3 * m.presc_node[:, None] + np.arange(3)).ravel()] = True
    pl = presc[cols_local]
    A_loc[presc[rows], :] = 0.0
    A_loc[:, pl] = 0.0
    A_loc[np.arange(len(rows))[presc[rows]], np.arange(len(rows))[presc[rows]]] = 1.0

    # 3. distributed SpMV
    v_glob = rng.standard_normal(m.n_dof)
    v_loc = np.zeros((p.n_local, 3))
    v_loc[:p.n_own] = v_glob.reshape(-1, 3)[p.node_gid[:p.n_own]]
    _halo_exchange(dist, torch, p, v_loc)
    y_own = A_loc @ v_loc.ravel()

    # 4. Jacobi-PCG with all-reduced dots (the control flow of fea_gpu_solve)
    b_glob = rng.standard_normal(m.n_dof)
    b_glob[presc] = 0.0
    b = b_glob[rows]
    dinv = 1.0 / A_loc[np.arange(len(rows)), np.arange(len(rows))]
    n3 = len(rows)

    def allsum(*vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.numpy()

    u = np.zeros(n3); r = b.copy(); z = dinv * r
    pvec = np.zeros((p.n_local, 3)); pvec.ravel()[:n3] = z
    rz, bb = allsum(r @ z, b @ b)
    its = 0
    for its in range(1, 2000):
        _halo_exchange(dist, torch, p, pvec)
        q = A_loc @ pvec.ravel()
        (pq,) = allsum(pvec.ravel()[:n3] @ q)
        alpha = rz / pq
        u += alpha * pvec.ravel()[:n3]; r -= alpha * q
        z = dinv * r
        rz_new, rr = allsum(r @ z, r @ r)
        if rr <= 1e-24 * bb:
            break
        pvec.ravel()[:n3] = z + (rz_new / rz) * pvec.ravel()[:n3]
        rz = rz_new
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), rows=rows, y=y_own, u=u, its=its,
             counts=p.counts, v=v_glob, b=b_glob, x=x_glob)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(WORLD)]
    m = block_model((2, 6, 2), model=1, bc_style=1, dy=0.02)
    # the two ranks own every row exactly once
    rows = np.concatenate([p["rows"] for p in parts])
    assert sorted(rows) == list(range(m.n_dof))
    # reference: single-rank oracle matrix with the same clamping
    o = PortOracle(m)
    o.set_nodes(parts[0]["x"]); o.update_state(); o.assemble_stiffness()
    rp, ci, v = o.get_csr()
    K = np.zeros((m.n_dof, m.n_dof))
    K[np.repeat(np.arange(m.n_dof), np.diff(rp)), ci] = v
    presc = np.zeros(m.n_dof, bool)
    presc[(3 * m.presc_node[:, None] + np.arange(3)).ravel()] = True
    K[presc, :] = 0.0; K[:, presc] = 0.0; K[presc, presc] = 1.0
    y = np.zeros(m.n_dof); u = np.zeros(m.n_dof)
    for p in parts:
        y[p["rows"]] = p["y"]; u[p["rows"]] = p["u"]
    y_ref = K @ parts[0]["v"]
    assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    u_ref = np.linalg.solve(K, parts[0]["b"])
    assert np.abs(u - u_ref).max() <= 1e-8 * np.abs(u_ref).max()
    assert all(int(p["counts"][5]) == 1 for p in parts)          # slab partition: one neighbour each
    assert int(parts[0]["counts"][6]) == int(parts[1]["counts"][7])   # what 0 sends is what 1 receives
