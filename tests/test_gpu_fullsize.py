"""Cross-checks at the benchmark size (BASELINE.json configs[1]: 250 x 64 x 64 = 1,024,000 LSpace elements, 3.17 M equations,
250.8 M non-zeros) and of every alternate code path behind an environment switch.

The oracle cannot run at this size in test time; instead the independent GPU algorithms for the same quantity are compared
with each other value for value: cluster assembly (assemble_cluster.cu) vs owner-computes gather (assemble_gather.cu) vs
slot-map scatter with atomics (element_kernels.cu), blocked SpMV vs CSR SpMV, cooperative CG iteration vs the four-launch one.
Each of them is compared with the oracle / the reference at small sizes in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from conftest import relerr
from oofem_b200 import capi, meshgen
from oofem_b200.elements import ElementSet
from oofem_b200.linsolver import CudaCG
from oofem_b200.sparsemtrx import CudaCSR

pytestmark = pytest.mark.gpu

FULL = (250, 64, 64)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def beam():
    import torch
    nx, ny, nz = FULL
    coords, conn = meshgen.hex_beam(nx, ny, nz)
    mask = np.zeros((coords.shape[0], 3), bool)
    mask[:(ny + 1) * (nz + 1)] = True
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a, device=dev)
    return dict(coords=t(coords), conn=t(conn), loc=t(loc), matid=t(np.zeros(conn.shape[0], np.int32)), neq=neq, dev=dev,
                mat=np.array([[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64))


def _assemble(ctx, beam, A, path, monkeypatch, names):
    if path == "cluster":
        monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    else:
        monkeypatch.setenv("OB200_ASSEMBLY", path)
    S = ElementSet(ctx, "lspace", beam["coords"], beam["conn"], beam["matid"], beam["mat"], beam["loc"], beam["neq"])
    ctx.profile_reset()
    ctx.set_profiling(True)
    A.zero()
    S.assembleStiffness(A)
    ctx.sync()
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(k.startswith(names) for k in prof), (path, sorted(prof))
    v = A.values(device=True).clone()
    S.close()
    return v


def test_three_assembly_algorithms_agree_at_1m_hex(ctx, beam, monkeypatch):
    import torch
    A = CudaCSR(ctx)
    A.buildInternalStructure(beam["loc"], beam["neq"])
    assert A.giveNumberOfNonzeros() == 250760268
    vc = _assemble(ctx, beam, A, "cluster", monkeypatch, "lspace_cluster_kernel")
    vc2 = _assemble(ctx, beam, A, "cluster", monkeypatch, "lspace_cluster_kernel")
    assert torch.equal(vc, vc2), "cluster assembly is not bit-reproducible run to run"
    vg = _assemble(ctx, beam, A, "gather", monkeypatch, "lspace_gather_kernel")
    scale = float(vg.abs().max())
    assert float((vc - vg).abs().max()) / scale < 1e-13
    del vc2
    vs = _assemble(ctx, beam, A, "slotmap", monkeypatch, "lspace_stiffness_kernel")
    assert float((vc - vs).abs().max()) / scale < 1e-13
    assert float((vg - vs).abs().max()) / scale < 1e-13


def test_strip_assembly_agrees_with_cluster_assembly_at_1m_hex(ctx, beam, monkeypatch):
    """The general-tangent path (element matrices on the FP64 tensor path + rows from strips, assemble_strips.cu -- the one a
    MisesMat set takes) forced onto the linear benchmark mesh (OB200_ASSEMBLY=strips) against the cluster kernel."""
    import torch
    A = CudaCSR(ctx)
    A.buildInternalStructure(beam["loc"], beam["neq"])
    vc = _assemble(ctx, beam, A, "cluster", monkeypatch, "lspace_cluster_kernel")
    vs = _assemble(ctx, beam, A, "strips", monkeypatch, "lspace_rows_kernel")
    scale = float(vc.abs().max())
    assert float((vc - vs).abs().max()) / scale < 1e-13
    vs2 = _assemble(ctx, beam, A, "strips", monkeypatch, "lspace_ke_dmma_kernel")
    assert torch.equal(vs, vs2), "strip assembly is not bit-reproducible run to run"


def test_blocked_spmv_vs_csr_spmv_and_cg_variants_at_1m_hex(ctx, beam, monkeypatch):
    import torch
    monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    S = ElementSet(ctx, "lspace", beam["coords"], beam["conn"], beam["matid"], beam["mat"], beam["loc"], beam["neq"])
    A = CudaCSR(ctx)
    A.buildInternalStructure(beam["loc"], beam["neq"])
    S.assembleStiffness(A)
    monkeypatch.setenv("OB200_SPMV_BLOCKED", "0")
    B = CudaCSR(ctx)                                   # same matrix, CSR SpMV (the index is chosen when first needed)
    B.buildInternalStructure(beam["loc"], beam["neq"])
    B.set_values(A.values(device=True))
    g = torch.Generator(device=beam["dev"]).manual_seed(5)
    x = torch.randn(beam["neq"], dtype=torch.float64, device=beam["dev"], generator=g)
    ctx.profile_reset()
    ctx.set_profiling(True)
    yb = B.times(x)
    monkeypatch.delenv("OB200_SPMV_BLOCKED")
    ya = A.times(x)
    ctx.sync()
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(k.startswith("spmv_block_kernel") for k in prof) and any(k.startswith("spmv_stream_kernel") for k in prof), sorted(prof)
    assert float((ya - yb).abs().max()) / float(ya.abs().max()) < 1e-14
    # CG: cooperative tail vs four-launch iteration, blocked vs CSR product: same iterates to round-off after 30 iterations
    b = torch.ones(beam["neq"], dtype=torch.float64, device=beam["dev"])
    res = []
    for M, coop in ((A, "1"), (A, "0"), (B, "1")):
        monkeypatch.setenv("OB200_CG_COOP", coop)
        xs = torch.zeros(beam["neq"], dtype=torch.float64, device=beam["dev"])
        s = CudaCG(ctx).initializeFrom(dict(lstol=0.0, lsiter=30, lsprecond=1))
        s.solve(M, b, xs)
        ctx.sync()
        res.append((xs.clone(), s.last_iterations, s.last_residual))
    monkeypatch.delenv("OB200_CG_COOP")
    for xs, it, rs in res[1:]:
        assert it == res[0][1]
        assert float((xs - res[0][0]).abs().max()) / float(res[0][0].abs().max()) < 1e-10
        assert abs(rs - res[0][2]) < 1e-8 * abs(res[0][2])
    S.close()


def test_node_block_schedule_allpairs_cross_check(ctx, monkeypatch):
    """The all-pairs version of the node block schedule (kept as the cross-check of the sort-based one) yields the same
    assembled matrix through the gather kernel."""
    coords, conn = meshgen.hex_beam(20, 9, 7)
    coords = meshgen.perturb(coords, 0.02, seed=9)
    mask = np.zeros((coords.shape[0], 3), bool)
    mask[:(9 + 1) * (7 + 1)] = True
    mask[500, 1] = True                                # a node with one prescribed dof inside the mesh
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    mat = [[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]]
    vals = []
    for mode in ("allpairs", "sort"):
        monkeypatch.setenv("OB200_ASSEMBLY", "gather")
        if mode == "allpairs":
            monkeypatch.setenv("OB200_NODE_BLOCKS", "allpairs")
        else:
            monkeypatch.delenv("OB200_NODE_BLOCKS", raising=False)
        monkeypatch.setenv("OB200_SCHED_CACHE", "0")
        S = ElementSet(ctx, "lspace", coords, conn, np.zeros(conn.shape[0], np.int32), mat, loc, neq)
        A = CudaCSR(ctx)
        A.buildInternalStructure(loc, neq)
        ctx.profile_reset()
        ctx.set_profiling(True)
        S.assembleStiffness(A)
        ctx.sync()
        ctx.set_profiling(False)
        prof = ctx.profile_report()
        assert any(k.startswith("node_blocks_allpairs_kernel" if mode == "allpairs" else "node_blocks_kernel") for k in prof), sorted(prof)
        vals.append(A.values())
        S.close()
    assert relerr(vals[0], vals[1]) < 1e-15


def test_negative_volume_tetrahedron_is_an_error(ctx):
    """FEI3dTetLin::evaldNdx: OOFEM_ERROR("negative volume") for detJ <= 0 (/root/reference/src/core/fei3dtetlin.C:145-147)."""
    coords, conn = meshgen.tet_beam(3, 2, 2)
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], np.zeros((coords.shape[0], 3), bool))
    bad = conn.copy()
    bad[5, [1, 2]] = bad[5, [2, 1]]                    # swapping two vertices inverts the element
    with pytest.raises(capi.OofemB200Error) as ei:
        ElementSet(ctx, "ltrspace", coords, bad, np.zeros(conn.shape[0], np.int32), [[capi.MAT_ISOLE, 1e4, 0.3, 0, 0, 0, 0, 0]],
                   meshgen.location_arrays(bad, nodeeq), neq)
    assert "negative volume" in str(ei.value) and "element 6" in str(ei.value)
    ElementSet(ctx, "ltrspace", coords, conn, np.zeros(conn.shape[0], np.int32), [[capi.MAT_ISOLE, 1e4, 0.3, 0, 0, 0, 0, 0]],
               meshgen.location_arrays(conn, nodeeq), neq).close()


def test_times_transposed_and_is_allocated_at(ctx):
    """SparseMtrx::timesT (compcol.C:146-163) on an unsymmetric matrix; isAllocatedAt / at on an entry outside the pattern."""
    coords, conn = meshgen.hex_beam(5, 3, 3)
    mask = np.zeros((coords.shape[0], 3), bool)
    mask[:16] = True
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    A = CudaCSR(ctx)
    A.buildInternalStructure(loc, neq)
    rp, ci = A.structure()
    rng = np.random.default_rng(3)
    val = rng.normal(size=ci.size)
    A.set_values(val)
    x = rng.normal(size=neq)
    rows = np.repeat(np.arange(neq), np.diff(rp))
    yt = np.zeros(neq)
    np.add.at(yt, ci, val * x[rows])
    assert relerr(A.timesT(x), yt) < 1e-14
    y = np.zeros(neq)
    np.add.at(y, rows, val * x[ci])
    assert relerr(A.times(x), y) < 1e-14
    assert A.isAllocatedAt(1, 1) and A.at(1, 1) == val[0]
    far = neq                                          # last equation does not couple to the first
    assert not A.isAllocatedAt(1, far) and A.at(1, far) == 0.0
