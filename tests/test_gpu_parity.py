"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against
 (1) the reference-generated golden fixtures (tests/golden/*.npz), and
 (2) the CPU oracle on the same seeded inputs, at sizes the oracle finishes in seconds.

Tolerances (BASELINE.json north_star): CSR rowptr / colind bit-exact; element matrices and
internal forces 1e-12 relative (max-norm); converged displacements 1e-8 relative."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oofem_b200 import capi, meshgen
from oofem_b200.elements import ElementSet
from oofem_b200.engng import Domain, LinearStatic, StaticStructural
from oofem_b200.inputfile import DirichletBC, Material, NodalLoad, Problem
from oofem_b200.linsolver import CR_CONVERGED, CR_DIVERGED_ITS, CudaCG
from oofem_b200.sparsemtrx import CudaCSR
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_KE = 1e-12
TOL_U = 1e-8
LINEAR = ["lspace_cantilever", "lspace_prescribed", "ltrspace_cantilever"]
MISES = ["lspace_mises", "ltrspace_mises"]


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


# ---- against the reference's own output ------------------------------------------------

@pytest.mark.parametrize("keep", ["1", "0"])
@pytest.mark.parametrize("name", LINEAR + MISES)
def test_csr_structure_bit_exact_vs_reference(ctx, name, keep, monkeypatch):
    """keep = 1: the sorted rows of the count pass are kept and copied (one sort per row); 0: count pass + fill pass."""
    monkeypatch.setenv("OB200_PATTERN_KEEP", keep)
    pb, d = load_golden(name)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    ctx.profile_reset()
    ctx.set_profiling(True)
    A.buildInternalStructure(dom.loc, dom.neq)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(n.startswith("row_copy_kernel") for n in prof) == (keep == "1"), sorted(prof)
    rp, ci = A.structure()
    assert np.array_equal(rp, d["colptr"])
    assert np.array_equal(ci, d["rowind"])


@pytest.mark.parametrize("name", LINEAR)
def test_element_matrices_vs_reference(ctx, name):
    pb, d = load_golden(name)
    dom = Domain(ctx, pb)
    Ke = dom.elems.computeStiffnessMatrix()
    assert relerr(Ke.reshape(-1), d["elem_ke"]) < TOL_KE


@pytest.mark.parametrize("name", LINEAR)
def test_internal_forces_vs_reference(ctx, name):
    pb, d = load_golden(name)
    dom = Domain(ctx, pb)
    u = d["node_u"].reshape(-1, 3)
    fe = dom.elems.giveInternalForcesVector(u)
    assert relerr(fe.reshape(-1), d["elem_fint"]) < TOL_KE


@pytest.mark.parametrize("name", LINEAR)
def test_assembled_values_and_spmv_vs_reference(ctx, name):
    pb, d = load_golden(name)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    dom.elems.assembleStiffness(A)
    assert relerr(A.values(), d["val"]) < TOL_KE          # symmetric: CSR values == CompCol values
    y = A.times(d["spmv_x"])
    assert relerr(y, d["spmv_y"]) < TOL_KE
    # SparseMtrx::at, 1-based
    assert abs(A.at(1, 1) - d["val"][0]) <= TOL_KE * abs(d["val"]).max()
    # per-element assemble(loc, mat) path gives the same matrix
    B = CudaCSR(ctx)
    B.buildInternalStructure(dom.loc, dom.neq)
    Ke = d["elem_ke"].reshape(dom.loc.shape[0], dom.loc.shape[1], dom.loc.shape[1])
    B.assemble(dom.loc[0], Ke[0])
    B.assemble(dom.loc[1:], Ke[1:])
    assert relerr(B.values(), d["val"]) < 1e-14


@pytest.mark.parametrize("name", LINEAR)
def test_linear_static_displacements_vs_reference(ctx, name):
    pb, d = load_golden(name)
    em = LinearStatic(ctx, pb)
    u = em.solveYourselfAt(1.0)
    assert relerr(u, d["node_u"].reshape(-1, 3)) < TOL_U


@pytest.mark.parametrize("name", MISES)
def test_mises_newton_vs_reference(ctx, name):
    pb, d = load_golden(name)
    em = StaticStructural(ctx, pb)
    u = em.solveYourself()[-1]
    assert relerr(u, d["node_u"].reshape(-1, 3)) < TOL_U
    assert em.tangent_assemblies > len(em.iterations)      # tangent really re-assembled in the iterations
    st = em.domain.elems.state()
    assert (st[:, 6] > 0).any()                            # plastic flow happened
    fe = em.domain.elems.giveInternalForcesVector(d["node_u"].reshape(-1, 3))
    assert relerr(fe.reshape(-1), d["elem_fint"]) < TOL_U


# ---- against the oracle on seeded inputs -------------------------------------------------

def _random_problem(etype, nx, ny, nz, seed, mat):
    gen = meshgen.hex_beam if etype == "lspace" else meshgen.tet_beam
    lx = float(nx) / ny
    coords, conn = gen(nx, ny, nz, lx, 1.0, 1.0)
    fixed, tip = meshgen.cantilever_bcs(coords, lx)
    coords = meshgen.perturb(coords, 0.3 / max(nx, ny, nz) / 2, seed=seed)
    pb = Problem(engng="linearstatic", params=dict(nsteps=1, lstol=1e-13, lsiter=50000, lsprecond=1), coords=coords,
                 elem_type=etype, conn=conn, elem_mat=np.zeros(conn.shape[0], np.int32), materials=[mat])
    pb.ltfs[1] = ("const", 1.0)
    pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, fixed))
    pb.loads.append(NodalLoad([1, 2, 3], [0.1, 0.2, -1.0], 1, tip))
    return pb


@pytest.mark.parametrize("etype,dims", [("lspace", (12, 6, 5)), ("ltrspace", (8, 5, 4))])
def test_full_path_vs_oracle(ctx, etype, dims):
    pb = _random_problem(etype, *dims, seed=3, mat=Material("isole", 210e3, 0.3))
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    rp, ci = A.structure()
    assert np.array_equal(rp, md.colptr) and np.array_equal(ci, md.rowind)           # bit-exact
    Ke = dom.elems.computeStiffnessMatrix()
    Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams)
    assert relerr(Ke, Ke_o) < TOL_KE
    dom.elems.assembleStiffness(A)
    val_o = orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)
    assert relerr(A.values(), val_o) < TOL_KE
    rng = np.random.default_rng(0)
    x = rng.normal(size=dom.neq)
    assert relerr(A.times(x), orc.compcol_times(md.colptr, md.rowind, val_o, x)) < TOL_KE
    u = rng.normal(size=pb.coords.shape) * 1e-3
    fe = dom.elems.giveInternalForcesVector(u)
    fe_o = orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams,
                                     u[pb.conn - 1].reshape(pb.conn.shape[0], -1))
    assert relerr(fe, fe_o) < TOL_KE
    f = np.zeros(dom.neq)
    dom.elems.assembleInternalForces(u, f)
    assert relerr(f, orc.assemble_vector(md.loc, fe_o, md.neq)) < TOL_KE
    sol = orc.solve_linear_static(pb)
    ug = LinearStatic(ctx, pb).solveYourselfAt(1.0)
    assert relerr(ug, sol["u"]) < TOL_U


@pytest.mark.parametrize("path", ["cluster", "gather"])
def test_gather_assembly_irregular_numbering_vs_oracle(ctx, path, monkeypatch):
    """Both owner-computes kernels (OB200_ASSEMBLY=gather selects the round-1 kernel, the default is the cluster kernel
    of assemble_cluster.cu) on a mesh whose node and element numbers are random: a group of
    consecutive nodes then touches unrelated elements (more distinct elements than one geometry round
    holds, element lists of a node in arbitrary order), two materials, and a second assembly that
    accumulates on top of the first (SparseMtrx::assemble adds)."""
    if path == "gather":
        monkeypatch.setenv("OB200_ASSEMBLY", "gather")
    else:
        monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    pb = _random_problem("lspace", 9, 5, 4, seed=11, mat=Material("isole", 210e3, 0.3))
    rng = np.random.default_rng(7)
    nnode, nelem = pb.coords.shape[0], pb.conn.shape[0]
    perm = rng.permutation(nnode)                      # old node i -> new node perm[i]
    coords = np.empty_like(pb.coords)
    coords[perm] = pb.coords
    conn = (perm[pb.conn - 1] + 1).astype(np.int32)[rng.permutation(nelem)]
    pb.coords, pb.conn = coords, np.ascontiguousarray(conn)
    for bc in pb.bcs:
        bc.nodes = perm[bc.nodes - 1] + 1
    for ld in pb.loads:
        ld.nodes = perm[ld.nodes - 1] + 1
    pb.materials = [Material("isole", 210e3, 0.3), Material("isole", 70e3, 0.2)]
    pb.elem_mat = rng.integers(0, 2, size=nelem).astype(np.int32)
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    rp, ci = A.structure()
    assert np.array_equal(rp, md.colptr) and np.array_equal(ci, md.rowind)
    Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams)
    val_o = orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)
    ctx.profile_reset()
    ctx.set_profiling(True)
    A.zero()
    dom.elems.assembleStiffness(A)
    assert relerr(A.values(), val_o) < TOL_KE
    dom.elems.assembleStiffness(A)                     # no zero() in between: contributions add up
    assert relerr(A.values(), 2.0 * val_o) < TOL_KE
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    kname = "lspace_cluster_kernel" if path == "cluster" else "lspace_gather_kernel"
    assert prof.get(kname + "< false >", (0, 0))[1] == 1 and prof.get(kname + "< true >", (0, 0))[1] == 1, prof
    sol = orc.solve_linear_static(pb)
    ug = LinearStatic(ctx, pb).solveYourselfAt(1.0)
    assert relerr(ug, sol["u"]) < TOL_U


@pytest.mark.parametrize("etype", ["lspace", "ltrspace"])
def test_mises_material_point_vs_oracle(ctx, etype):
    """Stress return, state update, algorithmic tangent and commit on random strains."""
    mat = Material("misesmat", 210e3, 0.3, sig0=240.0, H=2100.0, omega_crit=0.2, a=30.0)
    pb = _random_problem(etype, 5, 3, 3, seed=5, mat=mat)
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    rng = np.random.default_rng(1)
    for scale in (2e-3, 6e-3, 4e-3):
        u = rng.normal(size=pb.coords.shape) * scale
        ue = u[pb.conn - 1].reshape(pb.conn.shape[0], -1)
        fe, eps, sig = dom.elems.giveInternalForcesVector(u, want_gp=True)
        fe_o, eps_o, sig_o = orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, ue,
                                                       md.state, want_gp=True)
        assert relerr(eps, eps_o) < TOL_KE and relerr(sig, sig_o) < TOL_KE and relerr(fe, fe_o) < TOL_KE
        st = dom.elems.state()
        assert (st[:, 14] > st[:, 6]).any()                   # some points are yielding
        assert relerr(st, md.state) < TOL_KE
        Ke = dom.elems.computeStiffnessMatrix()               # unsymmetric algorithmic tangent (damage on)
        Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, md.state)
        assert relerr(Ke, Ke_o) < TOL_KE
        assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() > 0.0
        dom.elems.updateYourself()
        orc.mises_commit(md.state)
    # unsymmetric element matrices land transposed-correctly in the row storage (one more
    # plastic increment, NOT committed: after the commit dKappa = 0 and the tangent is elastic)
    u = rng.normal(size=pb.coords.shape) * 8e-3
    dom.elems.giveInternalForcesVector(u)
    orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams,
                              u[pb.conn - 1].reshape(pb.conn.shape[0], -1), md.state)
    Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, md.state)
    assert np.abs(Ke_o - Ke_o.transpose(0, 2, 1)).max() > 0.0
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    dom.elems.assembleStiffness(A)
    val_cc = orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)       # column storage
    x = rng.normal(size=dom.neq)
    assert relerr(A.times(x), orc.compcol_times(md.colptr, md.rowind, val_cc, x)) < TOL_KE


@pytest.mark.parametrize("etype,mat,kernels,tet_rows", [
    ("lspace", "mises", ("lspace_ke_dmma_kernel", "lspace_rows_kernel"), None),
    ("ltrspace", "mises", ("tet_tangent_kernel", "ltrspace_rows2_kernel"), None),
    ("ltrspace", "isole", ("ltrspace_rows2_kernel",), None),
    ("ltrspace", "mises", ("tet_tangent_kernel", "ltrspace_rows_kernel"), "search"),       # the kernel for meshes beyond the tables' capacity
    ("ltrspace", "isole", ("ltrspace_rows_kernel",), "search"),
])
def test_owner_computes_general_tangent(ctx, etype, mat, kernels, tet_rows, monkeypatch):
    """MisesMat tangents (LSpace: element strips, assemble_strips.cu; LTRSpace: node rows, assemble_tet.cu) and the linear
    LTRSpace assemble without atomics and without the slot map: the kernels that ran are the owner-computes ones, the values
    agree with the oracle and with the slot-map path on the same material state, a second assembly accumulates, and the
    result is bit-identical run to run.  Irregular node / element numbering, two materials."""
    monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    if tet_rows:
        monkeypatch.setenv("OB200_TET_ROWS", tet_rows)
    else:
        monkeypatch.delenv("OB200_TET_ROWS", raising=False)
    m0 = (Material("misesmat", 210e3, 0.3, sig0=240.0, H=2100.0, omega_crit=0.2, a=30.0) if mat == "mises"
          else Material("isole", 210e3, 0.3))
    pb = _random_problem(etype, 7, 4, 3, seed=21, mat=m0)
    rng = np.random.default_rng(17)
    nnode, nelem = pb.coords.shape[0], pb.conn.shape[0]
    perm = rng.permutation(nnode)
    coords = np.empty_like(pb.coords)
    coords[perm] = pb.coords
    pb.coords, pb.conn = coords, np.ascontiguousarray((perm[pb.conn - 1] + 1).astype(np.int32)[rng.permutation(nelem)])
    for bc in pb.bcs:
        bc.nodes = perm[bc.nodes - 1] + 1
    for ld in pb.loads:
        ld.nodes = perm[ld.nodes - 1] + 1
    pb.materials = [m0, Material("isole", 70e3, 0.2)]
    pb.elem_mat = rng.integers(0, 2, size=nelem).astype(np.int32)
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    if mat == "mises":                                       # a plastic, uncommitted increment: unsymmetric tangents
        u = rng.normal(size=pb.coords.shape) * 6e-3
        dom.elems.giveInternalForcesVector(u)
        orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams,
                                  u[pb.conn - 1].reshape(nelem, -1), md.state)
        Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, md.state)
        assert np.abs(Ke_o - Ke_o.transpose(0, 2, 1)).max() > 0.0
    else:
        Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams)
    assert relerr(dom.elems.computeStiffnessMatrix(), Ke_o) < TOL_KE
    val_o = orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)          # column storage of K = row storage of K^T
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    x = rng.normal(size=dom.neq)
    y_o = orc.compcol_times(md.colptr, md.rowind, val_o, x)
    ctx.profile_reset()
    ctx.set_profiling(True)
    A.zero()
    dom.elems.assembleStiffness(A)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    for k in kernels:
        assert any(n.startswith(k) for n in prof), (k, sorted(prof))
    assert not any(n.startswith(("slot_map_kernel", "lspace_stiffness_kernel", "ltrspace_stiffness_kernel")) for n in prof), sorted(prof)
    v1 = A.values().copy()
    assert relerr(A.times(x), y_o) < TOL_KE
    A.zero()
    dom.elems.assembleStiffness(A)
    assert np.array_equal(A.values(), v1), "owner-computes assembly is not bit-reproducible run to run"
    dom.elems.assembleStiffness(A)                           # no zero(): SparseMtrx::assemble adds
    assert relerr(A.values(), 2.0 * v1) < 1e-15
    # the generic slot-map / atomic path on the same state
    monkeypatch.setenv("OB200_ASSEMBLY", "slotmap")
    dom2 = Domain(ctx, pb)
    if mat == "mises":
        dom2.elems.setState(dom.elems.state())
    B = CudaCSR(ctx)
    B.buildInternalStructure(dom2.loc, dom2.neq)
    dom2.elems.assembleStiffness(B)
    assert relerr(B.values(), v1) < 1e-13


def test_lspace_element_matrix_kernels_agree(ctx, monkeypatch):
    """The FP64 tensor-path element-matrix kernel (assemble_strips.cu) against the lane-per-block kernel (OB200_KE=dfma)."""
    pb = _random_problem("lspace", 6, 4, 3, seed=4, mat=Material("isole", 210e3, 0.3))
    dom = Domain(ctx, pb)
    ctx.profile_reset()
    ctx.set_profiling(True)
    k1 = dom.elems.computeStiffnessMatrix()
    monkeypatch.setenv("OB200_KE", "dfma")
    k2 = dom.elems.computeStiffnessMatrix()
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(n.startswith("lspace_ke_dmma_kernel") for n in prof) and any(n.startswith("lspace_stiffness_kernel") for n in prof), sorted(prof)
    assert relerr(k1, k2) < 1e-14


def _irregular_hex_problem(mat, nsec=5, nlayers=2, nring=2):
    """Unstructured LSpace mesh with irregular nodes: `nsec` quadrilateral sectors around a centre line (5 -> the nodes on the line
    have 5 hexahedra per layer around them, 10 with two layers: more than the 8 of a structured mesh, i.e. more than one chunk of
    incidences in the strip assembly), `nring` rings of elements, extruded in z; bottom layer clamped."""
    pts = [(0.0, 0.0)]
    idx = {}
    # ring r (1..nring): points on the rays (P) and between the rays (Q); ring 0 is the centre
    for r in range(1, nring + 1):
        for k in range(nsec):
            a = 2 * np.pi * k / nsec
            idx[("P", r, k)] = len(pts); pts.append((r * np.cos(a), r * np.sin(a)))
            b = 2 * np.pi * (k + 0.5) / nsec
            idx[("Q", r, k)] = len(pts); pts.append((1.25 * r * np.cos(b), 1.25 * r * np.sin(b)))
    quads = []
    for k in range(nsec):
        k1 = (k + 1) % nsec
        quads.append((0, idx[("P", 1, k1)], idx[("Q", 1, k)], idx[("P", 1, k)]))                       # clockwise seen from +z
        for r in range(1, nring):
            # two quads per sector and ring: P_r,k - Q_r,k - Q_r+1,k - P_r+1,k   and   Q_r,k - P_r,k1 - P_r+1,k1 - Q_r+1,k
            quads.append((idx[("P", r, k)], idx[("Q", r, k)], idx[("Q", r + 1, k)], idx[("P", r + 1, k)]))
            quads.append((idx[("Q", r, k)], idx[("P", r, k1)], idx[("P", r + 1, k1)], idx[("Q", r + 1, k)]))
    n2 = len(pts)
    coords = np.array([(x, y, 0.8 * z) for z in range(nlayers + 1) for (x, y) in pts])
    conn = []
    for z in range(nlayers):
        for q in quads:
            conn.append([v + (z + 1) * n2 + 1 for v in q] + [v + z * n2 + 1 for v in q])
    conn = np.array(conn, dtype=np.int32)
    pb = Problem(engng="linearstatic", params=dict(nsteps=1, lstol=1e-13, lsiter=50000, lsprecond=1), coords=coords, elem_type="lspace",
                 conn=conn, elem_mat=np.zeros(conn.shape[0], np.int32), materials=[mat])
    pb.ltfs[1] = ("const", 1.0)
    pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, np.arange(1, n2 + 1)))
    pb.loads.append(NodalLoad([1, 2, 3], [0.1, 0.2, -1.0], 1, np.arange(nlayers * n2 + 1, (nlayers + 1) * n2 + 1)))
    return pb


@pytest.mark.parametrize("mat,path", [("isole", "cluster"), ("isole", "gather"), ("isole", "strips"), ("mises", None)])
def test_irregular_nodes_more_than_eight_elements_vs_oracle(ctx, mat, path, monkeypatch):
    """Nodes with 10 hexahedra around them (and a mix of 2, 4, 5, 8): every LSpace assembly kernel, the force kernels and the
    solve against the oracle."""
    if path in (None, "cluster"):
        monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    else:
        monkeypatch.setenv("OB200_ASSEMBLY", path)
    m = Material("isole", 210e3, 0.3) if mat == "isole" else Material("misesmat", 210e3, 0.3, sig0=240.0, H=2100.0, omega_crit=0.2, a=30.0)
    pb = _irregular_hex_problem(m)
    md = orc.Model(pb)
    nelem = pb.conn.shape[0]
    inc = np.bincount(pb.conn.reshape(-1) - 1)
    assert inc.max() == 10
    dom = Domain(ctx, pb)
    rng = np.random.default_rng(8)
    state = None
    u = rng.normal(size=pb.coords.shape) * (6e-3 if mat == "mises" else 1e-3)
    f, ebe = np.zeros(dom.neq), np.zeros(3)
    dom.elems.assembleInternalForces(u, f, ebe)
    fe_o = orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, u[pb.conn - 1].reshape(nelem, -1), md.state)
    assert relerr(f, orc.assemble_vector(md.loc, fe_o, md.neq)) < TOL_KE
    assert relerr(ebe, (fe_o.reshape(nelem, -1, 3) ** 2).sum(axis=(0, 1))) < TOL_KE
    if mat == "mises":
        state = md.state
    Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, state)
    assert relerr(dom.elems.computeStiffnessMatrix(), Ke_o) < TOL_KE
    val_o = orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    rp, ci = A.structure()
    assert np.array_equal(rp, md.colptr) and np.array_equal(ci, md.rowind)
    ctx.profile_reset()
    ctx.set_profiling(True)
    A.zero()
    dom.elems.assembleStiffness(A)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    want = {"cluster": "lspace_cluster_kernel", "gather": "lspace_gather_kernel", "strips": "lspace_rows_kernel", None: "lspace_rows_kernel"}[path]
    assert any(n.startswith(want) for n in prof), (want, sorted(prof))
    x = rng.normal(size=dom.neq)
    assert relerr(A.times(x), orc.compcol_times(md.colptr, md.rowind, val_o, x)) < TOL_KE
    dom.elems.assembleStiffness(A)
    assert relerr(A.times(x), 2.0 * orc.compcol_times(md.colptr, md.rowind, val_o, x)) < TOL_KE
    if mat == "isole":
        sol = orc.solve_linear_static(pb)
        assert relerr(LinearStatic(ctx, pb).solveYourselfAt(1.0), sol["u"]) < TOL_U


def test_tetrahedra_fan_beyond_the_table_capacity_vs_oracle(ctx, monkeypatch):
    """40 tetrahedra around one edge: its two nodes have 40 elements around them, more than the 32 the table-driven LTRSpace kernel
    holds -- tet_bind must fall back to the table-free kernel for the mesh (not to the slot map), with the same values."""
    monkeypatch.delenv("OB200_ASSEMBLY", raising=False)
    monkeypatch.delenv("OB200_TET_ROWS", raising=False)
    n = 40
    pts = [(0.0, 0.0, 0.0), (0.0, 0.0, 1.0)] + [(np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), 0.5) for k in range(n)]
    coords = np.array(pts)
    conn = np.array([[1, 2, 3 + k, 3 + (k + 1) % n] for k in range(n)], dtype=np.int32)
    pb = Problem(engng="linearstatic", params=dict(nsteps=1, lstol=1e-13, lsiter=50000, lsprecond=1), coords=coords, elem_type="ltrspace",
                 conn=conn, elem_mat=np.zeros(n, np.int32), materials=[Material("isole", 210e3, 0.3)])
    pb.ltfs[1] = ("const", 1.0)
    pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, np.array([3, 4, 5, 13, 23])))
    pb.loads.append(NodalLoad([1, 2, 3], [0.1, 0.2, -1.0], 1, np.array([2])))
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    ctx.profile_reset()
    ctx.set_profiling(True)
    dom.elems.assembleStiffness(A)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(k.startswith("ltrspace_rows_kernel") for k in prof) and not any(k.startswith(("ltrspace_rows2_kernel", "slot_map_kernel")) for k in prof), sorted(prof)
    val_o = orc.compcol_assemble(md.loc, orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams), md.colptr, md.rowind)
    assert relerr(A.values(), val_o) < TOL_KE
    sol = orc.solve_linear_static(pb)
    assert relerr(LinearStatic(ctx, pb).solveYourselfAt(1.0), sol["u"]) < TOL_U


@pytest.mark.parametrize("etype", ["lspace", "ltrspace"])
def test_owner_computes_internal_force_assembly(ctx, etype, monkeypatch):
    """EngngModel::assembleVector with InternalForceAssembler without atomics (node_force_gather_kernel): against the oracle
    (vector and element-by-element norms), against the atomic scatter (OB200_VECTOR_ASSEMBLY=atomic), accumulating into a
    non-zero vector, and bit-identical run to run.  Irregular numbering, MisesMat + IsoLE."""
    monkeypatch.delenv("OB200_VECTOR_ASSEMBLY", raising=False)
    m0 = Material("misesmat", 210e3, 0.3, sig0=240.0, H=2100.0, omega_crit=0.2, a=30.0)
    pb = _random_problem(etype, 7, 4, 3, seed=33, mat=m0)
    rng = np.random.default_rng(5)
    nnode, nelem = pb.coords.shape[0], pb.conn.shape[0]
    perm = rng.permutation(nnode)
    coords = np.empty_like(pb.coords)
    coords[perm] = pb.coords
    pb.coords, pb.conn = coords, np.ascontiguousarray((perm[pb.conn - 1] + 1).astype(np.int32)[rng.permutation(nelem)])
    for bc in pb.bcs:
        bc.nodes = perm[bc.nodes - 1] + 1
    for ld in pb.loads:
        ld.nodes = perm[ld.nodes - 1] + 1
    pb.materials = [m0, Material("isole", 70e3, 0.2)]
    pb.elem_mat = rng.integers(0, 2, size=nelem).astype(np.int32)
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    u = rng.normal(size=pb.coords.shape) * 5e-3
    fe_o = orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, u[pb.conn - 1].reshape(nelem, -1), md.state)
    f_o = orc.assemble_vector(md.loc, fe_o, md.neq)
    ebe_o = (fe_o.reshape(nelem, -1, 3) ** 2).sum(axis=(0, 1))
    ctx.profile_reset()
    ctx.set_profiling(True)
    f, ebe = np.zeros(dom.neq), np.zeros(3)
    dom.elems.assembleInternalForces(u, f, ebe)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    assert any(n.startswith("node_force_gather_kernel") for n in prof) and any(n.startswith("norm_finish_kernel") for n in prof), sorted(prof)
    assert relerr(f, f_o) < TOL_KE and relerr(ebe, ebe_o) < TOL_KE
    f2, ebe2 = np.zeros(dom.neq), np.zeros(3)
    dom.elems.assembleInternalForces(u, f2, ebe2)
    assert np.array_equal(f, f2) and np.array_equal(ebe, ebe2), "owner-computes vector assembly is not bit-reproducible"
    base = rng.normal(size=dom.neq)
    f3 = base.copy()
    dom.elems.assembleInternalForces(u, f3)                       # adds to what is there
    assert relerr(f3 - base, f_o) < 1e-10
    monkeypatch.setenv("OB200_VECTOR_ASSEMBLY", "atomic")
    f4, ebe4 = np.zeros(dom.neq), np.zeros(3)
    dom.elems.assembleInternalForces(u, f4, ebe4)
    assert relerr(f4, f) < 1e-13 and relerr(ebe4, ebe) < 1e-13


@pytest.mark.parametrize("etype", ["lspace", "ltrspace"])
@pytest.mark.parametrize("mat", ["isole", "mises"])
def test_extrapolated_forces_vs_oracle(ctx, etype, mat):
    """f += sum_e Ke du_e (StaticStructural::assembleExtrapolatedForces), evaluated as B^T (D B du) dV without forming Ke:
    against the oracle's element matrices -- for MisesMat the unsymmetric algorithmic tangent of a plastic, uncommitted state."""
    m = Material("isole", 70e3, 0.25) if mat == "isole" else Material("misesmat", 210e3, 0.3, sig0=240.0, H=2100.0, omega_crit=0.2, a=30.0)
    pb = _random_problem(etype, 6, 3, 3, seed=9, mat=m)
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    rng = np.random.default_rng(2)
    nelem = pb.conn.shape[0]
    state = None
    if mat == "mises":
        u = rng.normal(size=pb.coords.shape) * 6e-3
        dom.elems.giveInternalForcesVector(u)
        orc.batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, u[pb.conn - 1].reshape(nelem, -1), md.state)
        state = md.state
    du = rng.normal(size=pb.coords.shape) * 1e-3
    f = np.zeros(dom.neq)
    dom.elems.assembleExtrapolatedForces(du, f)
    Ke = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, state)
    if mat == "mises":
        assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() > 0.0
    due = du[pb.conn - 1].reshape(nelem, -1)
    ref = orc.assemble_vector(md.loc, np.einsum("eij,ej->ei", Ke, due), md.neq)
    assert relerr(f, ref) < TOL_KE
    f2 = np.zeros(dom.neq)
    dom.elems.assembleExtrapolatedForces(du, f2)
    assert np.array_equal(f, f2)


# ---- solver semantics and edge cases -----------------------------------------------------

def test_cg_iteration_semantics_match_iml(ctx):
    """Same iteration count and residual as the IML template at a loose tolerance; flag 1 and
    iters == max_iter when the budget is exhausted (iml/cg.h:70-71)."""
    pb, d = load_golden("lspace_cantilever")
    md = orc.Model(pb)
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc, dom.neq)
    dom.elems.assembleStiffness(A)
    b = np.ones(dom.neq)
    for precond in (0, 1):
        xo, flag_o, it_o, res_o = orc.cg(md.colptr, md.rowind, d["val"], b, precond=precond, max_iter=2000, tol=1e-6)
        s = CudaCG(ctx).initializeFrom(dict(lstol=1e-6, lsiter=2000, lsprecond=precond))
        x = np.zeros(dom.neq)
        assert s.solve(A, b, x) == CR_CONVERGED and flag_o == 0
        assert abs(s.last_iterations - it_o) <= 1
        assert relerr(x, xo) < 1e-5
    s = CudaCG(ctx).initializeFrom(dict(lstol=1e-14, lsiter=5, lsprecond=1))
    x = np.zeros(dom.neq)
    assert s.solve(A, b, x) == CR_DIVERGED_ITS and s.last_iterations == 5 and s.last_residual > 1e-14
    # initial guess already the solution: zero iterations (cg.h:36-41)
    s = CudaCG(ctx).initializeFrom(dict(lstol=1e-6, lsiter=100, lsprecond=1))
    x = xo.copy()
    assert s.solve(A, b, x) == CR_CONVERGED and s.last_iterations == 0
    # zero right-hand side: normb := 1, x stays 0
    x = np.zeros(dom.neq)
    assert s.solve(A, np.zeros(dom.neq), x) == CR_CONVERGED and s.last_iterations == 0 and not x.any()


def test_error_behaviour(ctx):
    pb, d = load_golden("ltrspace_cantilever")
    dom = Domain(ctx, pb)
    A = CudaCSR(ctx)
    A.buildInternalStructure(dom.loc[:10], dom.neq)            # structure from a subset of the elements
    with pytest.raises(capi.OofemB200Error) as ei:             # entries outside the structure
        dom.elems.assembleStiffness(A)
    assert ei.value.code == capi.ESTRUCT
    with pytest.raises(capi.OofemB200Error):                   # incompatible dimensions in times()
        A.times(np.zeros(dom.neq + 1))
    with pytest.raises(capi.OofemB200Error) as ei:             # at() out of bounds
        A.at(dom.neq + 1, 1)
    assert ei.value.code == capi.EINVAL
    Z = CudaCSR(ctx)
    Z.buildInternalStructure(dom.loc, dom.neq)                 # all-zero matrix: zero diagonal in DiagPreconditioner
    with pytest.raises(capi.OofemB200Error) as ei:
        CudaCG(ctx).initializeFrom(dict(lsprecond=1)).solve(Z, np.ones(dom.neq), np.zeros(dom.neq))
    assert ei.value.code == capi.EZERODIAG
    bad = dom.loc.copy()
    bad[0, 0] = dom.neq + 5
    with pytest.raises(capi.OofemB200Error):
        CudaCSR(ctx).buildInternalStructure(bad, dom.neq)


def test_empty_and_degenerate_inputs(ctx):
    A = CudaCSR(ctx)
    A.buildInternalStructure(np.zeros((0, 24), np.int32), 0)
    assert A.giveNumberOfRows() == 0 and A.giveNumberOfNonzeros() == 0
    # every dof prescribed: elements exist, no equations
    coords, conn = meshgen.hex_beam(2, 1, 1)
    loc = np.zeros((2, 24), np.int32)
    A.buildInternalStructure(loc, 0)
    assert A.giveNumberOfNonzeros() == 0
    S = ElementSet(ctx, "lspace", coords, conn, np.zeros(2, np.int32), [[1, 10.0, 0.2, 0, 0, 0, 0, 0]], loc, 0)
    S.assembleStiffness(A)
    Ke = S.computeStiffnessMatrix()
    assert np.isfinite(Ke).all() and np.allclose(Ke, Ke.transpose(0, 2, 1), atol=1e-12 * np.abs(Ke).max())
    # rigid translation produces no internal forces
    fe = S.giveInternalForcesVector(np.ones((coords.shape[0], 3)))
    assert np.abs(fe).max() < 1e-12 * np.abs(Ke).max()


@pytest.mark.parametrize("etype,dims", [("lspace", (7, 4, 3)), ("ltrspace", (5, 4, 3))])
def test_element_set_from_nodal_equation_numbers(ctx, etype, dims, monkeypatch):
    """ob200_elemset_create_nodal: the location arrays formed on the device from nodeeq[nnode][3] are those of
    Element::giveLocationArray -- tangent and internal forces assembled through such a set are bit-identical to the set
    created from the host's location arrays (host pointers and device-resident inputs); connectivity outside 1..nnode is
    an error."""
    import torch
    monkeypatch.setenv("OB200_SCHED_CACHE", "0")               # every set builds its own schedule
    pb = _random_problem(etype, *dims, seed=17, mat=Material("isole", 3.0e4, 0.25))
    dom = Domain(ctx, pb)
    u = np.random.default_rng(2).normal(size=pb.coords.shape) * 1e-3

    def through(S):
        A = CudaCSR(ctx)
        A.buildInternalStructure(dom.loc, dom.neq)
        S.assembleStiffness(A)
        f = np.zeros(dom.neq)
        S.assembleInternalForces(u, f)
        return A.values(), f

    v0, f0 = through(dom.elems)
    md = orc.Model(pb)
    Ke_o = orc.batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams)
    assert relerr(v0, orc.compcol_assemble(md.loc, Ke_o, md.colptr, md.rowind)) < TOL_KE
    S1 = ElementSet(ctx, etype, pb.coords, pb.conn, pb.elem_mat, pb.matparams(), None, dom.neq, nodeeq=dom.nodeeq)
    v1, f1 = through(S1)
    assert np.array_equal(v0, v1) and np.array_equal(f0, f1)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    S2 = ElementSet(ctx, etype, t(pb.coords), t(pb.conn), t(pb.elem_mat), pb.matparams(), None, dom.neq, nodeeq=t(dom.nodeeq.astype(np.int32)))
    torch.cuda.synchronize()
    v2, f2 = through(S2)
    assert np.array_equal(v0, v2) and np.array_equal(f0, f2)
    bad = pb.conn.copy()
    bad[1, 2] = pb.coords.shape[0] + 3
    with pytest.raises(capi.OofemB200Error) as ei:
        ElementSet(ctx, etype, pb.coords, bad, pb.elem_mat, pb.matparams(), None, dom.neq, nodeeq=dom.nodeeq)
    assert ei.value.code == capi.EINVAL
    with pytest.raises(capi.OofemB200Error):                   # both or neither of loc / nodeeq
        ElementSet(ctx, etype, pb.coords, pb.conn, pb.elem_mat, pb.matparams(), dom.loc, dom.neq, nodeeq=dom.nodeeq)


def test_device_resident_inputs(ctx):
    """on_device = 1: torch CUDA tensors are consumed in place (the `value` path of bench.py)."""
    import torch
    pb = _random_problem("lspace", 6, 4, 4, seed=11, mat=Material("isole", 1.0e4, 0.3))
    dom = Domain(ctx, pb)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(a, device=dev)
    S = ElementSet(ctx, "lspace", t(pb.coords), t(pb.conn), t(pb.elem_mat), pb.matparams(), t(dom.loc), dom.neq)
    torch.cuda.synchronize()
    A = CudaCSR(ctx)
    A.buildInternalStructure(t(dom.loc), dom.neq)
    S.assembleStiffness(A)
    B = CudaCSR(ctx)
    B.buildInternalStructure(dom.loc, dom.neq)
    dom.elems.assembleStiffness(B)
    assert np.array_equal(A.structure()[1], B.structure()[1])
    assert relerr(A.values(), B.values()) < 1e-13
    b = torch.ones(dom.neq, dtype=torch.float64, device=dev)
    x = torch.zeros(dom.neq, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    s = CudaCG(ctx).initializeFrom(dict(lstol=1e-10, lsiter=5000, lsprecond=1))
    assert s.solve(A, b, x) == CR_CONVERGED
    ctx.sync()
    xh = np.zeros(dom.neq)
    s.solve(B, np.ones(dom.neq), xh)
    assert relerr(x.cpu().numpy(), xh) < 1e-8


def test_linearity_and_symmetry_at_scale(ctx):
    """Size-independent properties on a mesh the oracle would take long for (200k elements):
    A symmetric (x.Ay == y.Ax), rigid-body translations in the null space of the free-free
    operator, CG solution satisfies the system."""
    nx, ny, nz = 128, 40, 40
    coords, conn = meshgen.hex_beam(nx, ny, nz)
    fixed, tip = meshgen.cantilever_bcs(coords, float(nx) / ny)
    mask = np.zeros((coords.shape[0], 3), bool)
    mask[fixed - 1] = True
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    S = ElementSet(ctx, "lspace", coords, conn, np.zeros(conn.shape[0], np.int32), [[1, 210e3, 0.3, 0, 0, 0, 0, 0]], loc, neq)
    A = CudaCSR(ctx)
    A.buildInternalStructure(loc, neq)
    S.assembleStiffness(A)
    rp, ci = A.structure()
    assert rp[-1] == A.giveNumberOfNonzeros() and np.all(np.diff(rp) > 0) and np.diff(rp).max() == 81
    rows = np.repeat(np.arange(neq), np.diff(rp))
    assert np.all(ci[1:][rows[1:] == rows[:-1]] > ci[:-1][rows[1:] == rows[:-1]])       # sorted rows
    rng = np.random.default_rng(4)
    x, y = rng.normal(size=neq), rng.normal(size=neq)
    Ax, Ay = A.times(x), A.times(y)
    assert abs(x @ Ay - y @ Ax) < 1e-10 * abs(x @ Ay)
    assert relerr(A.times(2.0 * x + y), 2.0 * Ax + Ay) < 1e-13
    # rigid translation of the whole body: internal forces vanish
    f = np.zeros(neq)
    S.assembleInternalForces(np.ones_like(coords), f)
    assert np.abs(f).max() < 1e-9 * np.abs(A.values()).max()
    b = rng.normal(size=neq)
    xs = np.zeros(neq)
    s = CudaCG(ctx).initializeFrom(dict(lstol=1e-10, lsiter=20000, lsprecond=1))
    assert s.solve(A, b, xs) == CR_CONVERGED
    assert np.linalg.norm(A.times(xs) - b) / np.linalg.norm(b) < 2e-10
