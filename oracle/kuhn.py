"""ORACLE / TEST INFRASTRUCTURE ONLY.

numpy generator of the synthetic Kuhn (Freudenthal) block used by bench.py's reference arm and parity
block, so that neither has to load the product library to obtain a mesh.  Same numbering as the
product's host mesher (fea_mesh_block): node (jx, jy, jz) on the half grid has id (jy * pz + jz) * px + jx;
each cube is cut into the six tets that share the body diagonal; local node order follows the
reference's shape functions (fea_solver.c:1287-1300: 0..3 vertices, 4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3)
8=(1,3) 9=(2,3)); vertices 1 and 2 are swapped where needed for a positive Jacobian (J = dN . x, :690-696).
tests/test_oracle_golden.py checks the two generators against each other.
"""
from __future__ import annotations

import itertools

import numpy as np

_PERMS = [p for p in itertools.permutations(range(3))]          # (0,1,2) (0,2,1) (1,0,2) (1,2,0) (2,0,1) (2,1,0)
_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]


def kuhn_block(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, y0=0.0, bc_style=0, dy=0.0, cube_origin=(0, 0, 0), full=None):
    """Nodes, connectivity and prescribed displacements of an nx x ny x nz block.

    With `full` = (NX, NY, NZ) and `cube_origin` = (ox, oy, oz) the block is the window of cubes
    [ox, ox+nx) x [oy, oy+ny) x [oz, oz+nz) of a larger NX x NY x NZ block: node coordinates are those of
    the large block and `gid` maps the window's nodes / elements to the large block's ids (no boundary
    conditions are produced for a window)."""
    px, py, pz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    jy, jz, jx = np.meshgrid(np.arange(py), np.arange(pz), np.arange(px), indexing="ij")
    if full is None:
        nodes = np.stack([lx * jx / (px - 1), y0 + ly * jy / (py - 1), lz * jz / (pz - 1)], axis=-1).reshape(-1, 3)
        node_gid = None
    else:
        NX, NY, NZ = full
        PX, PY, PZ = 2 * NX + 1, 2 * NY + 1, 2 * NZ + 1
        gx, gy, gz = jx + 2 * cube_origin[0], jy + 2 * cube_origin[1], jz + 2 * cube_origin[2]
        nodes = np.stack([lx * gx / (PX - 1), y0 + ly * gy / (PY - 1), lz * gz / (PZ - 1)], axis=-1).reshape(-1, 3)
        node_gid = ((gy * PZ + gz) * PX + gx).reshape(-1).astype(np.int64)
    nodes = np.ascontiguousarray(nodes, np.float64)

    def nid(vx, vy, vz):
        return (vy * pz + vz) * px + vx

    cy, cz, cx = np.meshgrid(np.arange(ny), np.arange(nz), np.arange(nx), indexing="ij")
    cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)
    ncube = cx.size
    conn = np.empty((ncube, 6, 10), np.int64)
    for t, perm in enumerate(_PERMS):
        v = np.zeros((4, 3), np.int64)
        for s in range(3):
            v[s + 1] = v[s]
            v[s + 1, perm[s]] += 2
        a = (v[1:] - v[0]).astype(float)
        if np.linalg.det(a) < 0:
            v[[1, 2]] = v[[2, 1]]
        off = [(v[k, 0], v[k, 1], v[k, 2]) for k in range(4)]
        off += [((v[i, 0] + v[j, 0]) // 2, (v[i, 1] + v[j, 1]) // 2, (v[i, 2] + v[j, 2]) // 2) for i, j in _EDGES]
        for k, (ox, oy, oz) in enumerate(off):
            conn[:, t, k] = nid(2 * cx + ox, 2 * cy + oy, 2 * cz + oz)
    conn = np.ascontiguousarray(conn.reshape(-1, 10), np.int32)
    out = dict(nodes=nodes, conn=conn)
    if full is not None:
        NX, NY, NZ = full
        cube_g = ((cy + cube_origin[1]) * NZ + (cz + cube_origin[2])) * NX + (cx + cube_origin[0])
        out["node_gid"] = node_gid
        out["elem_gid"] = (cube_g[:, None] * 6 + np.arange(6)[None, :]).reshape(-1).astype(np.int64)
        out["presc_node"] = np.zeros(0, np.int32)
        out["presc_type"] = np.zeros(0, np.int32)
        out["presc_vals"] = np.zeros((0, 3))
        return out
    pn, pt, pv = [], [], []
    for side in range(2):
        yy = py - 1 if side else 0
        zz, xx = np.meshgrid(np.arange(pz), np.arange(px), indexing="ij")
        ids = nid(xx, yy, zz).reshape(-1)
        typ = np.full(ids.size, 7 if bc_style == 1 else 2, np.int32)
        if bc_style != 1 and side == 0:
            typ[(xx.reshape(-1) == 0) & (zz.reshape(-1) == 0)] = 7
            if bc_style == 2:
                typ[(xx.reshape(-1) == px - 1) & (zz.reshape(-1) == 0)] = 6
        val = np.zeros((ids.size, 3))
        val[:, 1] = dy if side else 0.0
        pn.append(ids); pt.append(typ); pv.append(val)
    out["presc_node"] = np.concatenate(pn).astype(np.int32)
    out["presc_type"] = np.concatenate(pt).astype(np.int32)
    out["presc_vals"] = np.ascontiguousarray(np.concatenate(pv))
    return out
