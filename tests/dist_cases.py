"""Multi-GPU parity cases shared by tests/dist_worker.py (pytest, needs >= 2 GPUs) and by the checker leg of
bench.py (runs after the timed region whenever WORLD_SIZE > 1, so the driver's 2/4/8-GPU scaling runs carry
the evidence).  TEST INFRASTRUCTURE: imports the oracle.

A small perturbed beam (LSpace bricks or LTRSpace tetrahedra) is cut into WORLD_SIZE element partitions
(oofem2part-style node-cut partition, /root/reference/tools/oofem2part.py:137 classifyNodes), either x-slabs
(every shared dof has two sharers) or a px x py x pz box partition (dofs on the inner edges / at the centre are
shared by 4 / 8 ranks).  Every rank assembles its own elements on its GPU, shared dofs are completed by the
communicator (EngngModel::updateSharedDofManagers, /root/reference/src/core/engngm.C:2162) and
ob200_cg_solve_dist solves the global system.  Checked against the serial oracle on the unpartitioned mesh:
halo-completed SpMV 1e-12 relative, displacements 1e-8 relative (north_star tolerances)."""
import os

import numpy as np

SPMV_TOL = 1e-12
U_TOL = 1e-8


def box_factors(world):
    """world = px * py * pz with the factors as equal as possible (2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2)."""
    f = [1, 1, 1]
    w, k = world, 0
    for p in (2, 3, 5, 7):
        while w % p == 0:
            f[k % 3] *= p
            w //= p
            k += 1
    if w > 1:
        f[0] *= w
    return f


def element_partition(kind, world, nx, ny, nz, per_cell):
    """element -> rank map of a structured beam of (nx, ny, nz) cells, per_cell elements per cell
    (1 for bricks, 6 for the tetrahedral split), cells numbered x slowest like meshgen."""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    if kind == "slab":
        part = (ix * world) // nx
    else:
        px, py, pz = box_factors(world)
        part = (ix * px) // nx + px * ((iy * py) // ny + py * ((iz * pz) // nz))
    return np.repeat(part.reshape(-1), per_cell).astype(np.int32)


def run_case(ctx, dev, etype, kind, transport):
    """One parity case on the initialised torch.distributed group.  Returns a dict (same on all ranks apart
    from 'rank'); raises nothing: 'ok' says whether the tolerances held."""
    import torch
    import torch.distributed as dist
    from oofem_b200 import meshgen, partition
    from oofem_b200.comm import Comm
    from oofem_b200.elements import ElementSet
    from oofem_b200.linsolver import CudaCG
    from oofem_b200.sparsemtrx import CudaCSR
    from oracle import oracle as orc

    rank, world = dist.get_rank(), dist.get_world_size()
    px, py, pz = box_factors(world) if kind == "box" else (world, 1, 1)
    nx, ny, nz = 3 * px, max(3, 3 * py), max(3, 3 * pz)
    if etype == "lspace":
        coords_g, conn_g = meshgen.hex_beam(nx, ny, nz, nx / 3.0, ny / 3.0, nz / 3.0)
        per_cell, oet = 1, orc.LSPACE
    else:
        coords_g, conn_g = meshgen.tet_beam(nx, ny, nz, nx / 3.0, ny / 3.0, nz / 3.0)
        per_cell, oet = 6, orc.LTRSPACE
    coords_g = meshgen.perturb(coords_g, 0.03, seed=3)
    fixed_g = np.zeros((coords_g.shape[0], 3), bool)
    fixed_g[:(ny + 1) * (nz + 1)] = True                       # clamp x = 0
    matp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    # serial oracle on the whole mesh
    nodeeq_g, neq_g = meshgen.equation_numbers(coords_g.shape[0], fixed_g)
    loc_g = meshgen.location_arrays(conn_g, nodeeq_g)
    cp, ri = orc.compcol_build(loc_g, neq_g)
    val_g = orc.compcol_assemble(loc_g, orc.batch_stiffness(oet, conn_g, coords_g, np.zeros(conn_g.shape[0], np.int32), matp), cp, ri)
    rng = np.random.default_rng(2)
    bg, xg = rng.standard_normal(neq_g), rng.standard_normal(neq_g)
    xs = orc.cg(cp, ri, val_g, bg, precond=1, max_iter=5000, tol=1e-13)[0]
    yg = orc.compcol_times(cp, ri, val_g, xg)

    # this rank's partition on its GPU
    epart = element_partition(kind, world, nx, ny, nz, per_cell)
    part = partition.partition_mesh(coords_g, conn_g, epart, rank, world)
    sharers = np.ones(part.coords.shape[0], np.int32)
    for r, nodes in part.shared_nodes.items():
        sharers[nodes] += 1
    nodeeq, neq = meshgen.equation_numbers(part.coords.shape[0], fixed_g[part.node_global])
    loc = meshgen.location_arrays(part.conn, nodeeq)
    l2g = nodeeq_g[part.node_global].reshape(-1)[nodeeq.reshape(-1) > 0] - 1
    S = ElementSet(ctx, etype, part.coords, part.conn, np.zeros(part.conn.shape[0], np.int32), matp, loc, neq)
    A = CudaCSR(ctx)
    A.buildInternalStructure(loc, neq)
    A.zero()
    S.assembleStiffness(A)
    old = os.environ.get("OB200_P2P")
    os.environ["OB200_P2P"] = "1" if transport == "p2p" else "0"
    try:
        comm = Comm.from_torch_distributed(ctx, dev).set_halo(neq, *partition.halo_arrays(part, nodeeq, neq))
    finally:
        if old is None:
            os.environ.pop("OB200_P2P", None)
        else:
            os.environ["OB200_P2P"] = old

    y = A.times(torch.as_tensor(xg[l2g], device=dev))
    ctx.sync()
    comm.exchange_add(y)
    ctx.sync()
    e_spmv = float(np.abs(y.cpu().numpy() - yg[l2g]).max() / np.abs(yg).max())

    solver = CudaCG(ctx, comm).initializeFrom(dict(lstol=1e-13, lsiter=5000, lsprecond=1))
    x = np.zeros(neq)
    flag = solver.solve(A, np.ascontiguousarray(bg[l2g]), x)
    e_u = float(np.abs(x - xs[l2g]).max() / np.abs(xs).max())
    v = torch.tensor([e_spmv, e_u, float(solver.last_iterations), -float(solver.last_iterations), float(flag), float(sharers.max()),
                      1.0 if comm.p2p else 0.0, 0.0 if comm.p2p else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    v = v.cpu().numpy()
    used_p2p = bool(v[6] > 0.5) and not bool(v[7] > 0.5)
    ok = bool(v[0] < SPMV_TOL and v[1] < U_TOL and v[4] == 0 and v[2] == -v[3] and used_p2p == (transport == "p2p"))
    dist.barrier()           # nobody unmaps a mailbox a peer may still write
    comm.close()
    S.close()
    dist.barrier()
    return {"etype": etype, "partition": kind if kind == "slab" else "box %dx%dx%d" % (px, py, pz), "transport": transport,
            "world": world, "ok": ok, "spmv_relerr": float(v[0]), "u_relerr": float(v[1]), "iters": int(v[2]),
            "iters_equal_on_all_ranks": bool(v[2] == -v[3]), "flag": int(v[4]), "max_sharers": int(v[5]),
            "transport_used": "p2p" if used_p2p else "nccl", "neq_global": int(neq_g), "rank": rank}


def run_all(ctx, dev, cases=None):
    """Every combination the verdict of round 1 asked for: LSpace and LTRSpace, slabs and boxes, both transports."""
    out = []
    if cases is None:
        cases = [(e, k, t) for e in ("lspace", "ltrspace") for k in ("slab", "box") for t in ("p2p", "nccl")]
    for e, k, t in cases:
        out.append(run_case(ctx, dev, e, k, t))
    return out
