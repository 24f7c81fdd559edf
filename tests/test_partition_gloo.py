"""N > 1 host logic on CPU: two gloo ranks, each holding one element partition.

What the product's multi-GPU path does on the device (comm.cu: pack shared equations, exchange with
the neighbour, add in ascending rank order; cg.cu: owned-only dot products + all-reduce) is replayed
here with torch.distributed/gloo on the halo description produced by oofem_b200.partition -- the
arrays that go verbatim into ob200_comm_set_halo.  The local operators come from the oracle (the
checker).  Compared against the serial oracle on the unpartitioned mesh.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oofem_b200 import meshgen, partition  # noqa: E402

NX, NY, NZ = 3, 3, 2          # per-rank slab


def _global_problem(world):
    coords, conn = meshgen.hex_beam(world * NX, NY, NZ, world * NX / NY, 1.0, NZ / NY)
    coords = meshgen.perturb(coords, 0.04, seed=5)
    fixed = np.zeros((coords.shape[0], 3), bool)
    fixed[:(NY + 1) * (NZ + 1)] = True
    fixed[-1, 1] = True                                  # a single prescribed dof inside the last slab
    return coords, conn, fixed


def _local_system(part, coords_g, fixed_g):
    from oracle import oracle as orc
    fixed = fixed_g[part.node_global]
    nodeeq, neq = meshgen.equation_numbers(part.coords.shape[0], fixed)
    loc = meshgen.location_arrays(part.conn, nodeeq)
    colptr, rowind = orc.compcol_build(loc, neq)
    matp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    Ke = orc.batch_stiffness(orc.LSPACE, part.conn, part.coords, np.zeros(part.conn.shape[0], np.int32), matp)
    val = orc.compcol_assemble(loc, Ke, colptr, rowind)
    return nodeeq, neq, (colptr, rowind, val)


def _exchange_add(y, rank, neigh, offs, eqs):
    """comm_exchange_add (oofem_b200/csrc/comm.cu): own + neighbours' values, ascending rank order."""
    send = y[eqs].copy()
    recv = np.zeros_like(send)
    reqs = []
    for k, r in enumerate(neigh):
        s = torch.from_numpy(send[offs[k]:offs[k + 1]].copy())
        t = torch.zeros(offs[k + 1] - offs[k], dtype=torch.float64)
        reqs.append((dist.isend(s, int(r)), dist.irecv(t, int(r)), t, k))
    for a, b, t, k in reqs:
        a.wait(); b.wait()
        recv[offs[k]:offs[k + 1]] = t.numpy()
    contrib = {}
    for k, r in enumerate(neigh):
        for t in range(offs[k], offs[k + 1]):
            contrib.setdefault(int(eqs[t]), []).append((int(r), recv[t]))
    out = y.copy()
    for e, lst in contrib.items():
        lst.append((rank, y[e]))
        acc = 0.0
        for _, v in sorted(lst):
            acc += v
        out[e] = acc
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        coords_g, conn_g, fixed_g = _global_problem(world)
        epart = (np.arange(conn_g.shape[0]) // (NX * NY * NZ)).astype(np.int32)
        part = partition.partition_mesh(coords_g, conn_g, epart, rank, world)
        # the locally built slab is the same partition
        slab = partition.slab_partition(NX, NY, NZ, rank, world)
        assert np.array_equal(slab.conn, part.conn) and np.array_equal(slab.node_global, part.node_global)
        assert np.array_equal(slab.node_owner, part.node_owner)
        assert sorted(slab.shared_nodes) == sorted(part.shared_nodes)
        for r in part.shared_nodes:
            assert np.array_equal(slab.shared_nodes[r], part.shared_nodes[r])
        nodeeq, neq, (colptr, rowind, val) = _local_system(part, coords_g, fixed_g)
        neigh, offs, eqs, owned = partition.halo_arrays(part, nodeeq, neq)

        # serial reference on the unpartitioned mesh (every rank computes it)
        nodeeq_g, neq_g = meshgen.equation_numbers(coords_g.shape[0], fixed_g)
        loc_g = meshgen.location_arrays(conn_g, nodeeq_g)
        cp_g, ri_g = orc.compcol_build(loc_g, neq_g)
        matp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
        Ke_g = orc.batch_stiffness(orc.LSPACE, conn_g, coords_g, np.zeros(conn_g.shape[0], np.int32), matp)
        val_g = orc.compcol_assemble(loc_g, Ke_g, cp_g, ri_g)
        l2g = nodeeq_g[part.node_global].reshape(-1)[nodeeq.reshape(-1) > 0] - 1      # local eq -> global eq

        # 1. every global equation is owned by exactly one rank
        cnt = np.zeros(neq_g)
        np.add.at(cnt, l2g[owned == 1], 1.0)
        tc = torch.from_numpy(cnt)
        dist.all_reduce(tc)
        assert np.array_equal(tc.numpy(), np.ones(neq_g))

        # 2. distributed SpMV = local SpMV + halo exchange-add
        rng = np.random.default_rng(11)
        xg = rng.standard_normal(neq_g)
        y = _exchange_add(orc.compcol_times(colptr, rowind, val, xg[l2g]), rank, neigh, offs, eqs)
        yg = orc.compcol_times(cp_g, ri_g, val_g, xg)
        assert np.abs(y - yg[l2g]).max() <= 1e-12 * np.abs(yg).max()

        # 3. distributed PCG (IML++ CG statement by statement, cg.cu's distributed form) = serial CG
        def ddot(a, b):
            t = torch.tensor([float(np.dot(a[owned == 1], b[owned == 1]))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t[0])
        bg = rng.standard_normal(neq_g)
        b = bg[l2g]
        diag = np.zeros(neq)
        for j in range(neq):
            for p in range(colptr[j], colptr[j + 1]):
                if rowind[p] == j:
                    diag[j] = val[p]
        diag = _exchange_add(diag, rank, neigh, offs, eqs)
        x = np.zeros(neq)
        r = b - _exchange_add(orc.compcol_times(colptr, rowind, val, x), rank, neigh, offs, eqs)
        normb = np.sqrt(ddot(b, b))
        rho_1 = 0.0
        p = np.zeros(neq)
        iters = 0
        for it in range(1, 2001):
            z = r / diag
            rho = ddot(r, z)
            p = z.copy() if it == 1 else z + (rho / rho_1) * p
            qv = _exchange_add(orc.compcol_times(colptr, rowind, val, p), rank, neigh, offs, eqs)
            alpha = rho / ddot(p, qv)
            x += alpha * p
            r -= alpha * qv
            iters = it
            if np.sqrt(ddot(r, r)) / normb <= 1e-12:
                break
            rho_1 = rho
        ref = orc.cg(cp_g, ri_g, val_g, bg, precond=1, max_iter=2000, tol=1e-12)
        xs = ref["x"] if isinstance(ref, dict) else ref[0]
        err = np.abs(x - xs[l2g]).max() / np.abs(xs).max()
        assert err < 1e-8, err
        q.put((rank, "ok", iters))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partition_halo_and_cg_gloo():
    from oracle import oracle as orc
    orc.lib()                                            # build once, before forking
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    assert res[0][2] == res[1][2]                        # both ranks took the same number of iterations


def test_halo_arrays_shapes_single_rank():
    part = partition.slab_partition(2, 2, 2, 0, 1)
    nodeeq, neq = meshgen.equation_numbers(part.coords.shape[0], np.zeros((part.coords.shape[0], 3), bool))
    neigh, offs, eqs, owned = partition.halo_arrays(part, nodeeq, neq)
    assert neigh.size == 0 and offs.tolist() == [0] and eqs.size == 0 and owned.all()


def test_tet_slab_partition_matches_partition_of_global_mesh():
    """bench.py --etype ltrspace builds every rank's slab of tetrahedra locally: identical to cutting the global tet_beam by
    the cell's x-index (BASELINE configs[2], element partition a la oofem2part)."""
    nx, ny, nz, world = 3, 2, 2, 3
    h = 1.0 / ny
    coords_g, conn_g = meshgen.tet_beam(world * nx, ny, nz, world * nx * h, ny * h, nz * h)
    per = 6 * nx * ny * nz
    # tet_beam numbers the six tetrahedra of a cell consecutively?  derive the owner from the element centroid instead
    cx = coords_g[conn_g - 1][:, :, 0].mean(axis=1)
    epart = np.minimum((cx / (nx * h)).astype(np.int32), world - 1)
    for rank in range(world):
        part = partition.partition_mesh(coords_g, conn_g, epart, rank, world)
        slab = partition.slab_partition(nx, ny, nz, rank, world, etype="ltrspace")
        assert slab.conn.shape == (per, 4) and np.array_equal(slab.node_global, part.node_global)
        assert np.allclose(slab.coords, part.coords, atol=1e-14)
        key = lambda c: np.sort(np.sort(c, axis=1).view([("", c.dtype)] * 4).ravel())
        assert np.array_equal(key(slab.conn), key(part.conn))           # the same tetrahedra (the local order may differ)
        assert np.array_equal(slab.node_owner, part.node_owner)
        assert sorted(slab.shared_nodes) == sorted(part.shared_nodes)
        for r in part.shared_nodes:
            assert np.array_equal(slab.shared_nodes[r], part.shared_nodes[r])
