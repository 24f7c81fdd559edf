"""Synthetic structured meshes for the BASELINE.json configurations.

Node ordering inside an element follows the reference's interpolation classes:
  * LSpace  -- FEI3dHexaLin::evalN (src/core/fei3dhexalin.C:46-63): nodes 1-4 on the
    zeta=+1 face ordered (-,-) (-,+) (+,+) (+,-) in (xi,eta), nodes 5-8 below them.
  * LTRSpace -- FEI3dTetLin (src/core/fei3dtetlin.C:141-147): positive detJ required
    ("negative volume" error at fei3dtetlin.C:145).
All numbers are 1-based like OOFEM's input records.
"""
from __future__ import annotations

import numpy as np


def beam_nodes(nx: int, ny: int, nz: int, lx: float, ly: float, lz: float) -> np.ndarray:
    """(nx+1)(ny+1)(nz+1) nodes, x slowest so that a slab of x is a contiguous node range."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    zs = np.linspace(0.0, lz, nz + 1)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float64)


def _nid(i, j, k, ny, nz):
    return (i * (ny + 1) + j) * (nz + 1) + k + 1


def hex_beam(nx: int, ny: int, nz: int, lx: float = None, ly: float = 1.0, lz: float = 1.0):
    """Structured LSpace beam along x.  Returns (coords[nnode,3] f64, conn[nelem,8] i32)."""
    if lx is None:
        lx = float(nx) * ly / ny
    coords = beam_nodes(nx, ny, nz, lx, ly, lz)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    n = lambda a, b, c: _nid(i + a, j + b, k + c, ny, nz)
    # (xi,eta,zeta) = (x,y,z): top face z+1 first
    conn = np.stack([n(0, 0, 1), n(0, 1, 1), n(1, 1, 1), n(1, 0, 1),
                     n(0, 0, 0), n(0, 1, 0), n(1, 1, 0), n(1, 0, 0)], axis=1).astype(np.int32)
    return coords, conn


def tet_beam(nx: int, ny: int, nz: int, lx: float = None, ly: float = 1.0, lz: float = 1.0):
    """Structured LTRSpace beam: every hex cell split into 6 tetrahedra around the main
    diagonal (conforming across cells).  Returns (coords, conn[nelem,4] i32)."""
    if lx is None:
        lx = float(nx) * ly / ny
    coords = beam_nodes(nx, ny, nz, lx, ly, lz)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    v = {}
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                v[(a, b, c)] = _nid(i + a, j + b, k + c, ny, nz)
    # Kuhn triangulation: the 6 monotone paths from (0,0,0) to (1,1,1)
    paths = [((1, 0, 0), (1, 1, 0)), ((1, 0, 0), (1, 0, 1)), ((0, 1, 0), (1, 1, 0)),
             ((0, 1, 0), (0, 1, 1)), ((0, 0, 1), (1, 0, 1)), ((0, 0, 1), (0, 1, 1))]
    tets = []
    for p1, p2 in paths:
        t = np.stack([v[(0, 0, 0)], v[p1], v[p2], v[(1, 1, 1)]], axis=1)
        tets.append(t)
    conn = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    # enforce positive detJ as defined by FEI3dTetLin (swap two nodes where negative)
    c = coords[conn - 1]
    x1, x2, x3, x4 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    d2, d3, d4 = x2 - x1, x3 - x1, x4 - x1
    det = (d4[:, 0] * d2[:, 1] * d3[:, 2] - d4[:, 0] * d3[:, 1] * d2[:, 2]
           + d3[:, 0] * d4[:, 1] * d2[:, 2] - d2[:, 0] * d4[:, 1] * d3[:, 2]
           + d2[:, 0] * d3[:, 1] * d4[:, 2] - d3[:, 0] * d2[:, 1] * d4[:, 2])
    neg = det < 0
    conn[neg, 1], conn[neg, 2] = conn[neg, 2].copy(), conn[neg, 1].copy()
    return coords, conn


def perturb(coords: np.ndarray, amp: float, seed: int = 0, keep_x0: bool = True) -> np.ndarray:
    """Deterministic interior jitter so that elements are not all congruent (parity tests)."""
    rng = np.random.default_rng(seed)
    d = rng.uniform(-amp, amp, size=coords.shape)
    out = coords + d
    if keep_x0:
        out[coords[:, 0] == 0.0] = coords[coords[:, 0] == 0.0]
    return out


def cantilever_bcs(coords: np.ndarray, lx: float):
    """Clamp the x=0 face; return (fixed_nodes, tip_nodes) as 1-based arrays."""
    fixed = np.nonzero(np.abs(coords[:, 0]) < 1e-12)[0] + 1
    tip = np.nonzero(np.abs(coords[:, 0] - lx) < 1e-9 * max(1.0, lx))[0] + 1
    return fixed.astype(np.int32), tip.astype(np.int32)


def equation_numbers(nnode: int, fixed_mask: np.ndarray):
    """Default OOFEM numbering (EngngModel::forceEquationNumbering, src/core/engngm.C:483-486):
    nodes in order, dofs in order, consecutive numbers for free dofs, 0 for prescribed.
    fixed_mask[nnode,3] bool.  Returns (nodeeq[nnode,3] int32 1-based, neq)."""
    free = ~fixed_mask.reshape(-1)
    eq = np.zeros(nnode * 3, dtype=np.int32)
    eq[free] = np.arange(1, int(free.sum()) + 1, dtype=np.int32)
    return eq.reshape(nnode, 3), int(free.sum())


def location_arrays(conn: np.ndarray, nodeeq: np.ndarray) -> np.ndarray:
    """Element::giveLocationArray for 3-dof nodes: loc[e, 3a+d] = eq(node a, dof d)."""
    return nodeeq[conn - 1].reshape(conn.shape[0], -1).astype(np.int32)
