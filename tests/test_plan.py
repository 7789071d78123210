"""CPU: host-side partition + symbolic phase (fea_plan_*), checked against the oracle's
assembled matrix and by simulating the gather map with numpy."""
import numpy as np
import pytest

import fea_gpu as fg
from conftest import block_model, csr_mv, load_golden
from oracle.oracle import PortOracle



def ke_code(a, b):
    """Staging order of the a<=b blocks of K_e, restated from fea_plan.hpp: region pr = min(a, 9-a)
    holds rows pr and 9-pr as five consecutive pairs, the left-over block (a,9) of the odd row last."""
    pr = min(a, 9 - a)
    first = list(range(pr, 10))            # columns of row pr
    second = list(range(9 - pr, 10))       # columns of row 9-pr
    order = []
    for row, cols in ((pr, first), (9 - pr, second)):
        order += [(row, c) for c in cols[:len(cols) // 2 * 2]]
    order += [(row, 9) for row, cols in ((pr, first), (9 - pr, second)) if len(cols) % 2]
    assert len(order) == 11
    return 11 * pr + order.index((a, b))


TRI = {(a, b): ke_code(a, b) for a in range(10) for b in range(a, 10)}
assert sorted(TRI.values()) == list(range(55))


def test_staging_layout_invariants_the_kernels_rely_on():
    """K_e staging (fea_plan.hpp): block idx = 55 e + code sits at double offset 9 idx + idx // 11 =
    500 e + 100 pr + 9 pos.  The element kernel stores block pairs with 16-byte stores and the gather
    reads the aligned 80-byte window around a block with 16-byte loads: pairs must start on even
    offsets, blocks must not overlap, and the window must stay inside the element's 500 doubles."""
    used = np.zeros(500, dtype=int)
    for (a, b), code in TRI.items():
        pr, pos = divmod(code, 11)
        assert pr == min(a, 9 - a)
        for e in (0, 1, 7, 123456):
            idx = 55 * e + code
            assert 9 * idx + idx // 11 == 500 * e + 100 * pr + 9 * pos
        off = 100 * pr + 9 * pos
        used[off:off + 9] += 1
        if pos % 2 == 0:                                   # first block of a pair, or the single at pos 10
            assert off % 2 == 0
        if pos < 10 and pos % 2 == 0:                      # its partner follows immediately
            partner = [ab for ab, c in TRI.items() if c == code + 1]
            assert partner == [(a, b + 1)]
        lo = off & ~1
        assert 0 <= lo and lo + 10 <= 500                  # five double2 loads stay inside the element
    assert used.max() == 1 and used.sum() == 495           # 55 disjoint blocks, 5 pad doubles
    assert [i for i in range(500) if not used[i]] == [99, 199, 299, 399, 499]


def staged_blocks(o, n_elems):
    """The K_e staging buffer as element_kernel lays it out: [e][500] doubles, block (a<=b) at
    100 pr + 9 pos (fea_plan.hpp), pad doubles poisoned so that a wrong offset shows."""
    out = np.full((n_elems, 500), np.nan)
    for e in range(n_elems):
        ke = o.element_matrix(e).reshape(10, 3, 10, 3)
        for (a, b), code in TRI.items():
            off = 100 * (code // 11) + 9 * (code % 11)
            out[e, off:off + 9] = ke[a, :, b, :].ravel()
    return out


def gather_numpy(plan, staged):
    """What gather_blocks_kernel computes, in numpy: block idx = 55 e + code is read at double offset
    9 idx + idx // 11 of the flat staging buffer."""
    flat = staged.ravel()
    idx = (plan.csrc & 0x7fffffff).astype(np.int64)
    tr = (plan.csrc >> 31).astype(bool)
    off = 9 * idx + idx // 11
    blocks = flat[off[:, None] + np.arange(9)].reshape(-1, 3, 3)
    assert np.isfinite(blocks).all()
    blocks[tr] = blocks[tr].transpose(0, 2, 1)
    vals = np.zeros((plan.nnzb, 3, 3))
    np.add.at(vals, np.repeat(np.arange(plan.nnzb), np.diff(plan.cptr)), blocks)
    return vals


def test_interleaved_staging_variant_addresses_the_same_blocks():
    """-DFEA_KE_INTERLEAVED=1 (fea_plan.hpp): chunk c (16 bytes) of element e at double offset
    ((e // 32) * 250 + c) * 64 + 2 * (e % 32).  Writing K_e the way that element kernel does and reading
    the five chunks around a block the way that gather does must give the same matrix as the flat layout."""
    m = block_model((3, 2, 2))
    o = PortOracle(m); o.update_state(); o.assemble_stiffness()
    p = fg.Plan(m.nodes, m.conn)
    ne = p.n_elems
    flat = staged_blocks(o, len(m.conn))[p.elem_gid]                   # [e][500], local element order
    buf = np.full(((ne + 31) // 32) * 32 * 500, np.nan)
    e = np.arange(ne)
    for c in range(250):                                               # the element kernel's stores
        base = ((e // 32) * 250 + c) * 64 + 2 * (e % 32)
        buf[base], buf[base + 1] = flat[:, 2 * c], flat[:, 2 * c + 1]
    idx = (p.csrc & 0x7fffffff).astype(np.int64)
    el, code = idx // 55, idx % 55
    off = 100 * (code // 11) + 9 * (code % 11)
    w = np.empty((len(idx), 10))
    for h in range(5):                                                 # the gather's five 16-byte loads
        base = ((el // 32) * 250 + off // 2 + h) * 64 + 2 * (el % 32)
        w[:, 2 * h], w[:, 2 * h + 1] = buf[base], buf[base + 1]
    blocks = np.where((off % 2 == 1)[:, None], w[:, 1:10], w[:, 0:9]).reshape(-1, 3, 3)
    assert np.isfinite(blocks).all()
    tr = (p.csrc >> 31).astype(bool)
    blocks[tr] = blocks[tr].transpose(0, 2, 1)
    vals = np.zeros((p.nnzb, 3, 3))
    np.add.at(vals, np.repeat(np.arange(p.nnzb), np.diff(p.cptr)), blocks)
    assert np.array_equal(vals, gather_numpy(p, flat))


def bsr_to_dense_rows(plan, vals, n_dof):
    A = np.zeros((3 * plan.n_own, n_dof))
    for I in range(plan.n_own):
        for k in range(plan.browptr[I], plan.browptr[I + 1]):
            J = plan.node_gid[plan.bcol[k]]
            A[3 * I:3 * I + 3, 3 * J:3 * J + 3] = vals[k]
    return A


def test_pattern_and_gather_map_reproduce_oracle_matrix():
    m, z = load_golden("neohook_brick")
    o = PortOracle(m)
    o.apply_increment(1.0); o.update_state(); o.assemble_stiffness()
    rp, ci, v = o.get_csr()
    p = fg.Plan(m.nodes, m.conn)
    assert p.nnzb * 9 == len(v) == 145737
    assert sorted(p.node_gid) == list(range(len(m.nodes)))          # a permutation (Morton order)
    for I in range(p.n_own):   # full 3x3 blocks, identical node sets per row, ascending local columns
        g = p.node_gid[I]
        cols = p.bcol[p.browptr[I]:p.browptr[I + 1]]
        assert np.all(np.diff(cols) > 0)
        assert np.array_equal(np.sort(p.node_gid[cols]), ci[rp[3 * g]:rp[3 * g + 1]][::3] // 3)
    vals = gather_numpy(p, staged_blocks(o, len(m.conn))[p.elem_gid])   # csrc indexes LOCAL elements
    A = bsr_to_dense_rows(p, vals, m.n_dof)
    Aref = np.zeros((m.n_dof, m.n_dof))
    Aref[np.repeat(np.arange(m.n_dof), np.diff(rp)), ci] = v
    rows = (3 * p.node_gid[:p.n_own, None] + np.arange(3)).ravel()
    assert np.abs(A - Aref[rows]).max() <= 1e-13 * np.abs(Aref).max()
    # contributions are in ascending GLOBAL element id within each nonzero (the reference's order)
    e_of = p.elem_gid[(p.csrc & 0x7fffffff) // 55]
    for k in range(0, p.nnzb, 97):
        seg = e_of[p.cptr[k]:p.cptr[k + 1]]
        assert np.all(np.diff(seg.astype(np.int64)) >= 0)


def test_sell_layout_is_a_faithful_copy_of_the_block_csr():
    m = block_model((4, 5, 3))
    p = fg.Plan(m.nodes, m.conn)
    assert p.n_slices == (p.n_own + 31) // 32 and p.slice_ptr[-1] == p.n_slots
    assert np.all(np.diff(p.slice_ptr) % 32 == 0)
    live = p.sell_row[p.sell_row >= 0]
    assert sorted(live) == list(range(p.n_own))                  # every row sits in exactly one lane
    lens = np.diff(p.browptr)
    seen = 0
    for s in range(p.n_slices):
        width = (p.slice_ptr[s + 1] - p.slice_ptr[s]) // 32
        rows = p.sell_row[32 * s:32 * s + 32]
        assert width == max(lens[r] for r in rows if r >= 0)
        for lane, r in enumerate(rows):
            slots = p.slice_ptr[s] + 32 * np.arange(width) + lane
            if r < 0:
                assert np.all(np.diff(p.scptr[np.r_[slots, slots[-1] + 1]]) >= 0)
                continue
            n = lens[r]
            assert np.array_equal(p.sbcol[slots[:n]], p.bcol[p.browptr[r]:p.browptr[r + 1]])
            assert np.all(p.sbcol[slots[n:]] == r)               # padding points at the row itself
            for j in range(n):
                a, b = p.scptr[slots[j]], p.scptr[slots[j] + 1]
                q = p.browptr[r] + j
                assert np.array_equal(p.scsrc[a:b], p.csrc[p.cptr[q]:p.cptr[q + 1]])
                seen += b - a
            assert all(p.scptr[t] == p.scptr[t + 1] for t in slots[n:])
            jd = list(p.bcol[p.browptr[r]:p.browptr[r + 1]]).index(r)
            assert p.sdiag[r] == 9 * (p.slice_ptr[s] + 32 * jd) + lane
    assert seen == p.n_contrib
    assert p.n_slots / p.nnzb < 1.15                             # sigma-sorting keeps padding small


def test_kuhn_block_counts_and_geometry():
    n = 6
    mb = fg.mesh_block(n, n, n)
    assert mb["nodes"].shape == ((2 * n + 1) ** 3, 3) and mb["conn"].shape == (6 * n ** 3, 10)
    assert len(np.unique(mb["conn"])) == (2 * n + 1) ** 3          # every half-grid point is used
    X = mb["nodes"][mb["conn"]]
    for k, (a, b) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):   # fea_solver.c:1295-1300
        assert np.abs(X[:, 4 + k] - (X[:, a] + X[:, b]) / 2).max() < 1e-15
    m = block_model(n)
    o = PortOracle(m); o.update_state()
    g, detJ = o.get_gradients()
    assert detJ.min() > 0 and np.isclose((detJ * PortOracle.tables(5)[0][:, 0]).sum(), 1.0, rtol=1e-12)
    p = fg.Plan(m.nodes, m.conn)
    assert np.diff(p.browptr).max() * 3 == 195                     # max row of Kuhn meshes
    assert 2.2 < p.n_contrib / p.nnzb < 2.7


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_partition_halo_lists_are_consistent(nranks):
    m = block_model((3, 8, 3))
    plans = [fg.Plan(m.nodes, m.conn, r, nranks) for r in range(nranks)]
    owner = plans[0].owner
    assert sum(p.n_own for p in plans) == len(m.nodes)
    assert np.abs(np.bincount(owner, minlength=nranks) - len(m.nodes) / nranks).max() <= 1
    f = lambda gid: 1000.0 + 3.0 * gid                      # noqa: E731  value carried by node gid
    for r, p in enumerate(plans):
        assert np.all(owner[p.node_gid[:p.n_own]] == r) and len(set(p.node_gid[:p.n_own])) == p.n_own
        # every element touching an owned node is local, with all its nodes present
        touch = np.where((owner[m.conn] == r).any(axis=1))[0]
        assert np.array_equal(touch, np.sort(p.elem_gid))
        assert set(np.unique(m.conn[touch])) == set(p.node_gid)
        for i, q in enumerate(p.nbr_rank):
            pq = plans[q]
            j = list(pq.nbr_rank).index(r)
            sent = p.node_gid[p.send_nodes[p.send_ptr[i]:p.send_ptr[i + 1]]]
            ghosts = pq.node_gid[pq.n_own + pq.recv_ptr[j]: pq.n_own + pq.recv_ptr[j + 1]]
            assert np.array_equal(sent, ghosts)               # same nodes, same order on both sides
            assert np.array_equal(f(sent), f(ghosts))
        assert p.recv_ptr[-1] == p.n_local - p.n_own


def test_distributed_rows_reassemble_global_matrix():
    m = block_model((2, 4, 2), model=0)
    o = PortOracle(m)
    rng = np.random.default_rng(3)
    o.set_nodes(m.nodes + 0.01 * rng.standard_normal(m.nodes.shape))
    o.update_state(); o.assemble_stiffness()
    rp, ci, v = o.get_csr()
    x = rng.standard_normal(m.n_dof)
    y_ref = csr_mv(rp, ci, v, x)
    y = np.zeros(m.n_dof)
    for r in range(3):
        p = fg.Plan(m.nodes, m.conn, r, 3)
        oo = PortOracle(type(m)(**{**m.__dict__, "conn": np.ascontiguousarray(m.conn[p.elem_gid])}))
        oo.set_nodes(o.get_nodes()); oo.update_state()
        vals = gather_numpy(p, staged_blocks(oo, p.n_elems))   # element ids are local here
        A = bsr_to_dense_rows(p, vals, m.n_dof)
        rows = (3 * p.node_gid[:p.n_own, None] + np.arange(3)).ravel()
        y[rows] = A @ x
    assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max()


def test_bad_meshes_are_rejected():
    mb = fg.mesh_block(1, 1, 1)
    conn = mb["conn"].copy(); conn[0, 0] = 10 ** 6
    with pytest.raises(fg.FeaGpuError) as e:
        fg.Plan(mb["nodes"], conn)
    assert e.value.code == fg.ERR_MESH


def test_rcb_partition_boxes_and_odd_rank_counts():
    """Recursive coordinate bisection: a cube on 8 ranks becomes 2 x 2 x 2 boxes with up to 7
    neighbours each, a bar becomes slabs, and rank counts that are not powers of two still balance."""
    mb = fg.mesh_block(6, 6, 6, 6.0, 6.0, 6.0)
    nodes, conn = mb["nodes"], mb["conn"]
    p0 = fg.Plan(nodes, conn, 0, 8)
    own = p0.owner
    counts = np.bincount(own, minlength=8)
    assert counts.min() >= len(nodes) // 8 - 1 and counts.max() <= len(nodes) // 8 + 1
    for r in range(8):
        ext = np.ptp(nodes[own == r], axis=0)
        assert np.all(ext <= 3.5), (r, ext)                     # half the cube in every direction
    # every rank agrees on the halo: what r sends to q is what q receives from r
    plans = [p0] + [fg.Plan(nodes, conn, r, 8) for r in range(1, 8)]
    nbrs = [p.n_nbr for p in plans]       # face, edge and corner neighbours (Kuhn tets couple along one body diagonal only)
    assert min(nbrs) >= 3 and max(nbrs) == 7, nbrs
    for r, p in enumerate(plans):
        for k, q in enumerate(p.nbr_rank):
            sent = p.node_gid[p.send_nodes[p.send_ptr[k]:p.send_ptr[k + 1]]]
            pq = plans[q]
            kk = list(pq.nbr_rank).index(r)
            recv = pq.node_gid[pq.n_own + pq.recv_ptr[kk]:pq.n_own + pq.recv_ptr[kk + 1]]
            assert np.array_equal(sent, recv), (r, q)
    # a bar: slabs along its length, two neighbours at most; three ranks balance to one node
    mb = fg.mesh_block(2, 9, 2, 2.0, 9.0, 2.0)
    p = [fg.Plan(mb["nodes"], mb["conn"], r, 3) for r in range(3)]
    counts = np.bincount(p[0].owner, minlength=3)
    assert counts.max() - counts.min() <= 1
    assert [q.n_nbr for q in p] == [1, 2, 1]
    ymax = [mb["nodes"][p[0].owner == r, 1].max() for r in range(3)]
    assert ymax[0] <= ymax[1] <= ymax[2]


def test_numpy_kuhn_mesher_matches_the_product_mesher():
    """oracle/kuhn.py (used by bench.py's reference arm and parity window) numbers the block exactly as
    fea_mesh_block does, windows of a larger block included."""
    from oracle.kuhn import kuhn_block
    for args in [(3, 4, 2, 1.0, 2.0, 3.0, 1.0, 0, 0.01), (2, 2, 2, 1.0, 1.0, 1.0, 0.0, 1, 0.02), (3, 2, 4, 3.0, 2.0, 4.0, 0.0, 2, 0.5)]:
        a, b = fg.mesh_block(*args), kuhn_block(*args)
        for k in a:
            assert np.array_equal(a[k], b[k]), (args, k)
    a = fg.mesh_block(5, 7, 4, 5.0, 7.0, 4.0, 0.0, 1, 0.0)
    w = kuhn_block(2, 3, 2, 5.0, 7.0, 4.0, 0.0, cube_origin=(1, 2, 1), full=(5, 7, 4))
    assert np.array_equal(a["nodes"][w["node_gid"]], w["nodes"])
    assert np.array_equal(w["node_gid"][w["conn"]], a["conn"][w["elem_gid"]])


@pytest.mark.parametrize("nranks,rank", [(1, 0), (2, 0), (2, 1), (3, 1)])
def test_cell_layout_of_the_direct_assembly_reproduces_the_pull_gather(nranks, rank):
    """Direct (push) assembly, fea_plan.cpp "cell layout": the element kernel writes every staged block into the cell
    edest names (transposed when bit 31 is set), gather_cells_kernel adds a column's layers position by position
    and writes each upper slot and its mirror.  Emulated here with the kernels' own index arithmetic and compared
    bit for bit with the pull gather over scptr / scsrc: same contributions, same order, every real slot written
    exactly once, padding untouched."""
    mb = fg.mesh_block(3, 2, 3, bc_style=1, dy=0.01)
    p = fg.Plan(mb["nodes"], mb["conn"], rank=rank, nranks=nranks)
    rng = np.random.default_rng(5)
    ne, ne_pad = p.n_elems, (p.n_elems + 31) // 32 * 32
    staged = rng.standard_normal((ne, 55, 3, 3))
    for a in range(10):                                    # K_e is symmetric: so are its diagonal blocks
        c = TRI[(a, a)]
        staged[:, c] = staged[:, c] + staged[:, c].transpose(0, 2, 1)
    # pull gather in slot order, one add per contribution starting from zero (gather_item)
    pull = np.zeros((p.n_slots, 3, 3))
    for s in range(p.n_slots):
        for k in range(p.scptr[s], p.scptr[s + 1]):
            src = int(p.scsrc[k])
            blk = staged[(src & 0x7fffffff) // 55, (src & 0x7fffffff) % 55]
            pull[s] = pull[s] + (blk.T if src >> 31 else blk)
    real = np.diff(p.scptr) > 0
    # element kernel: cells
    assert p.edest.shape == (55, ne_pad) and (p.edest[:, ne:] == 0xffffffff).all()
    cells = np.full((p.n_cells, 10), np.nan)
    written = np.zeros(p.n_cells, int)
    for e in range(ne):
        for code in range(55):
            d = int(p.edest[code, e])
            if d == 0xffffffff:
                continue
            blk = staged[e, code]
            cells[d & 0x7fffffff, :9] = (blk.T if d >> 31 else blk).ravel()
            written[d & 0x7fffffff] += 1
    assert (written == 1).all()
    # gather_cells_kernel
    out = np.full((p.n_slots, 3, 3), np.nan)
    hits = np.zeros(p.n_slots, int)
    assert len(set(p.col_order.tolist())) == len(p.col_order) == p.n_cols_active
    assert (np.diff(p.col_ready[p.col_order]) >= 0).all() and (p.col_ready[p.col_order] >= 0).all()
    assert p.n_cols_active == int((p.col_ready >= 0).sum())
    for col in p.col_order:
        meta = p.cmeta[32 * col:32 * col + 32].astype(int)
        n, rk = meta & 2047, meta >> 11
        off = int(p.ccell[col])
        acc = np.zeros((32, 9))
        ready = -1
        for k in range(n.max()):
            m = int((n > k).sum())
            acc[:m] = acc[:m] + cells[off:off + m, :9]
            off += m
        assert off == p.ccell[col + 1]
        for lane in range(32):
            if n[lane] == 0:
                continue
            slot = 32 * col + lane
            v = acc[rk[lane]].reshape(3, 3)
            out[slot] = v
            hits[slot] += 1
            mp = int(p.cmirror[slot])
            if mp >= 0:
                mcol, ml = divmod(mp, 288)
                assert ml < 32
                out[32 * mcol + ml] = v.T
                hits[32 * mcol + ml] += 1
        srcs = np.concatenate([p.scsrc[p.scptr[32 * col + l]:p.scptr[32 * col + l + 1]] for l in range(32) if n[l]])
        assert p.col_ready[col] == int(((srcs & 0x7fffffff) // 55).max())
    assert (hits[real] == 1).all() and (hits[~real] == 0).all()
    assert np.array_equal(out[real], pull[real])
    # upper slots are exactly the real slots with column >= row; ranks of a column are a permutation prefix
    row_of_slot = np.full(p.n_slots, -1)
    for s in range(p.n_slices):
        base, width = p.slice_ptr[s], (p.slice_ptr[s + 1] - p.slice_ptr[s]) // 32
        for l in range(32):
            r = p.sell_row[32 * s + l]
            if r >= 0:
                row_of_slot[base + 32 * np.arange(width) + l] = r
    upper = real & (p.sbcol >= row_of_slot)
    assert np.array_equal((p.cmeta & 2047) > 0, upper)
    assert np.array_equal((p.cmeta & 2047)[upper], np.diff(p.scptr)[upper])
