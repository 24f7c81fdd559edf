"""Element partitions for the multi-GPU path (one process per GPU).

OOFEM's parallel mode is a node-cut partitioning (tools/oofem2part, src/core/domain.C
`dofmanager ... shared partitions n r1..rn`): every rank owns a set of elements, nodes on a
partition boundary are replicated on all ranks touching them ("shared" dof managers), each rank
numbers its own equations and assembles only its own elements; shared dofs are completed by
summing the sharers' contributions (EngngModel::updateSharedDofManagers, src/core/engngm.C).

This module produces, per rank, the local mesh and the halo description
`ob200_comm_set_halo` (include/oofem_b200.h) expects:
  neigh_rank[nneigh], neigh_offset[nneigh+1], shared_eq[...] (0-based local equations, ordered by
  ascending GLOBAL dof id so that both sides of a pair agree on the order), owned[neq].
The lowest rank sharing a dof owns it (counts it in dot products).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import meshgen


@dataclass
class LocalPartition:
    rank: int
    nranks: int
    coords: np.ndarray            # [nnode_local, 3]
    conn: np.ndarray              # [nelem_local, nen] 1-based local node numbers
    elem_global: np.ndarray       # [nelem_local] 0-based global element numbers
    node_global: np.ndarray       # [nnode_local] 0-based global node numbers (ascending)
    shared_nodes: dict = field(default_factory=dict)   # neighbour rank -> local node indices (0-based), ascending global id
    node_owner: np.ndarray = None  # [nnode_local] owning rank of every local node


def partition_mesh(coords: np.ndarray, conn: np.ndarray, elem_part: np.ndarray, rank: int, nranks: int) -> LocalPartition:
    """Cut the global mesh (coords, conn 1-based) by the element -> rank map `elem_part`."""
    elem_part = np.asarray(elem_part)
    mine = np.nonzero(elem_part == rank)[0]
    gconn = conn[mine] - 1
    node_global = np.unique(gconn)                                   # ascending global ids
    lconn = (np.searchsorted(node_global, gconn) + 1).astype(np.int32)
    # ranks touching each of my nodes: scan the elements of the other ranks that use one of my nodes
    nnode = coords.shape[0]
    is_mine = np.zeros(nnode, dtype=bool)
    is_mine[node_global] = True
    owner = np.full(node_global.size, rank, dtype=np.int32)
    shared = {}
    for r in range(nranks):
        if r == rank:
            continue
        other = np.unique(conn[elem_part == r] - 1)
        common = other[is_mine[other]]                               # ascending global ids
        if common.size:
            loc = np.searchsorted(node_global, common)
            shared[r] = loc.astype(np.int64)
            owner[loc] = np.minimum(owner[loc], r)
    return LocalPartition(rank, nranks, coords[node_global].copy(), lconn, mine, node_global, shared, owner)


def slab_partition(nx: int, ny: int, nz: int, rank: int, nranks: int, h: float = None, etype: str = "lspace") -> LocalPartition:
    """x-slab `rank` of a structured LSpace beam nranks*nx elements long, built locally (no global
    mesh is ever materialised -- this is what bench.py uses at 1M elements per GPU).  Identical to
    partition_mesh(hex_beam(nranks*nx, ny, nz), element x-index // nx, rank) -- tested."""
    if h is None:
        h = 1.0 / ny
    gen = meshgen.hex_beam if etype == "lspace" else meshgen.tet_beam       # tet_beam: six tetrahedra per cell, cell-major
    coords, conn = gen(nx, ny, nz, nx * h, ny * h, nz * h)
    coords[:, 0] += rank * nx * h
    plane = (ny + 1) * (nz + 1)
    nloc = coords.shape[0]
    node_global = np.arange(nloc, dtype=np.int64) + rank * nx * plane
    owner = np.full(nloc, rank, dtype=np.int32)
    shared = {}
    if rank > 0:
        shared[rank - 1] = np.arange(plane, dtype=np.int64)
        owner[:plane] = rank - 1
    if rank < nranks - 1:
        shared[rank + 1] = np.arange(nloc - plane, nloc, dtype=np.int64)
    elem_global = np.arange(conn.shape[0], dtype=np.int64) + rank * conn.shape[0]
    return LocalPartition(rank, nranks, coords, conn, elem_global, node_global, shared, owner)


def halo_arrays(part: LocalPartition, nodeeq: np.ndarray, neq: int):
    """(neigh_rank i32[n], neigh_offset i64[n+1], shared_eq i32[...], owned u8[neq]) for
    ob200_comm_set_halo from the local equation numbers nodeeq[nnode_local, ndof] (1-based, 0 = prescribed)."""
    neigh, offs, eqs = [], [0], []
    owned = np.ones(neq, dtype=np.uint8)
    for r in sorted(part.shared_nodes):
        e = nodeeq[part.shared_nodes[r]].reshape(-1)
        e = e[e > 0] - 1
        if e.size == 0:
            continue
        neigh.append(r)
        eqs.append(e)
        offs.append(offs[-1] + e.size)
    not_mine = part.node_owner != part.rank
    e = nodeeq[not_mine].reshape(-1)
    owned[e[e > 0] - 1] = 0
    eqs = np.concatenate(eqs).astype(np.int32) if eqs else np.zeros(0, np.int32)
    return np.array(neigh, np.int32), np.array(offs, np.int64), eqs, owned
