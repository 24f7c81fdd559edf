"""Worker of tests/test_gpu_dist.py: run under torch.distributed.run, one rank per GPU.

A perturbed LSpace cantilever is cut into WORLD_SIZE element slabs (oofem_b200.partition); every rank
assembles its own elements on its GPU, shared dofs are completed over NCCL, and ob200_cg_solve_dist
solves the global system.  Checked against the serial oracle on the unpartitioned mesh:
displacements 1e-8 relative (north_star), halo-completed SpMV 1e-12."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oofem_b200 import capi, meshgen, partition  # noqa: E402
from oofem_b200.comm import Comm  # noqa: E402
from oofem_b200.elements import ElementSet  # noqa: E402
from oofem_b200.linsolver import CudaCG  # noqa: E402
from oofem_b200.sparsemtrx import CudaCSR  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, nz = 4, 5, 4
    coords_g, conn_g = meshgen.hex_beam(world * nx, ny, nz, world * nx / ny, 1.0, nz / ny)
    coords_g = meshgen.perturb(coords_g, 0.03, seed=3)
    fixed_g = np.zeros((coords_g.shape[0], 3), bool)
    fixed_g[:(ny + 1) * (nz + 1)] = True
    matp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    # serial oracle
    nodeeq_g, neq_g = meshgen.equation_numbers(coords_g.shape[0], fixed_g)
    loc_g = meshgen.location_arrays(conn_g, nodeeq_g)
    cp, ri = orc.compcol_build(loc_g, neq_g)
    val_g = orc.compcol_assemble(loc_g, orc.batch_stiffness(orc.LSPACE, conn_g, coords_g, np.zeros(conn_g.shape[0], np.int32), matp), cp, ri)
    rng = np.random.default_rng(2)
    bg, xg = rng.standard_normal(neq_g), rng.standard_normal(neq_g)
    xs = orc.cg(cp, ri, val_g, bg, precond=1, max_iter=5000, tol=1e-13)[0]
    yg = orc.compcol_times(cp, ri, val_g, xg)

    # this rank's partition on its GPU
    epart = (np.arange(conn_g.shape[0]) // (nx * ny * nz)).astype(np.int32)
    part = partition.partition_mesh(coords_g, conn_g, epart, rank, world)
    nodeeq, neq = meshgen.equation_numbers(part.coords.shape[0], fixed_g[part.node_global])
    loc = meshgen.location_arrays(part.conn, nodeeq)
    l2g = nodeeq_g[part.node_global].reshape(-1)[nodeeq.reshape(-1) > 0] - 1
    ctx = capi.Context(local)
    S = ElementSet(ctx, "lspace", part.coords, part.conn, np.zeros(part.conn.shape[0], np.int32), matp, loc, neq)
    A = CudaCSR(ctx)
    A.buildInternalStructure(loc, neq)
    A.zero()
    S.assembleStiffness(A)
    comm = Comm.from_torch_distributed(ctx, dev).set_halo(neq, *partition.halo_arrays(part, nodeeq, neq))

    y = A.times(torch.as_tensor(xg[l2g], device=dev))
    ctx.sync()
    comm.exchange_add(y)
    ctx.sync()
    e_spmv = float(np.abs(y.cpu().numpy() - yg[l2g]).max() / np.abs(yg).max())

    solver = CudaCG(ctx, comm).initializeFrom(dict(lstol=1e-13, lsiter=5000, lsprecond=1))
    x = np.zeros(neq)
    flag = solver.solve(A, np.ascontiguousarray(bg[l2g]), x)
    e_u = float(np.abs(x - xs[l2g]).max() / np.abs(xs).max())
    its = torch.tensor([solver.last_iterations], device=dev)
    lo, hi = its.clone(), its.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = e_spmv < 1e-12 and e_u < 1e-8 and flag == 0 and int(lo) == int(hi)
    print(json.dumps({"rank": rank, "ok": bool(ok), "spmv_relerr": e_spmv, "u_relerr": e_u, "flag": flag,
                      "iters": solver.last_iterations, "p2p": comm.p2p}), flush=True)
    dist.barrier()           # nobody unmaps a mailbox a peer may still write
    comm.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
