"""ctypes front end of the CPU oracle (oracle/oofem_oracle.c) plus a plain-numpy
restatement of the reference's LinearStatic / Newton-Raphson drivers.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PINNED against the reference: see tests/test_oracle_pinning.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LSPACE, LTRSPACE = 1, 2
ETYPE = {"lspace": LSPACE, "ltrspace": LTRSPACE}
NEN = {LSPACE: 8, LTRSPACE: 4}
NGP = {LSPACE: 8, LTRSPACE: 1}


def build() -> str:
    so = os.path.join(HERE, "liboofem_oracle.so")
    src = os.path.join(HERE, "oofem_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_compcol_build.restype = C.c_int64
        _LIB.orc_mises_state_doubles.restype = C.c_int
    return _LIB


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def isole_D(E, nu):
    D = np.zeros(36)
    lib().orc_isole_D(C.c_double(E), C.c_double(nu), _p(D))
    return D.reshape(6, 6)


def mises_state(nelem, etype):
    n = lib().orc_mises_state_doubles()
    return np.zeros((nelem * NGP[etype], n))


def mises_commit(state):
    lib().orc_mises_commit(C.c_int64(state.shape[0]), _p(state))


def batch_stiffness(etype, conn, coords, matid, matparams, state=None):
    conn, coords, matid, matparams = _i32(conn), _f64(coords), _i32(matid), _f64(matparams)
    nd = 3 * NEN[etype]
    Ke = np.zeros((conn.shape[0], nd, nd))
    lib().orc_batch_stiffness(C.c_int(etype), C.c_int64(conn.shape[0]), _p(conn), _p(coords), _p(matid),
                              _p(matparams), _p(state), _p(Ke))
    return Ke


def batch_internal_forces(etype, conn, coords, matid, matparams, ue, state=None, want_gp=False):
    conn, coords, matid, matparams, ue = _i32(conn), _f64(coords), _i32(matid), _f64(matparams), _f64(ue)
    ne, nd = conn.shape[0], 3 * NEN[etype]
    fe = np.zeros((ne, nd))
    eps = np.zeros((ne * NGP[etype], 6)) if want_gp else None
    sig = np.zeros((ne * NGP[etype], 6)) if want_gp else None
    lib().orc_batch_internal_forces(C.c_int(etype), C.c_int64(ne), _p(conn), _p(coords), _p(matid), _p(matparams),
                                    _p(state), _p(ue), _p(fe), _p(eps), _p(sig))
    return (fe, eps, sig) if want_gp else fe


def compcol_build(loc, neq):
    loc = _i32(loc)
    colptr = np.zeros(neq + 1, dtype=np.int32)
    rp = C.c_void_p()
    nnz = lib().orc_compcol_build(C.c_int64(loc.shape[0]), C.c_int(loc.shape[1]), _p(loc), C.c_int32(neq),
                                  _p(colptr), C.byref(rp))
    rowind = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
    lib().orc_free(rp)
    return colptr, rowind


def compcol_assemble(loc, Ke, colptr, rowind, val=None):
    loc, Ke = _i32(loc), _f64(Ke)
    if val is None:
        val = np.zeros(rowind.shape[0])
    lib().orc_compcol_assemble(C.c_int64(loc.shape[0]), C.c_int(loc.shape[1]), _p(loc), _p(Ke), _p(colptr),
                               _p(rowind), _p(val))
    return val


def assemble_vector(loc, fe, neq):
    loc, fe = _i32(loc), _f64(fe)
    out = np.zeros(neq)
    lib().orc_assemble_vector(C.c_int64(loc.shape[0]), C.c_int(loc.shape[1]), _p(loc), _p(fe), _p(out))
    return out


def compcol_times(colptr, rowind, val, x):
    x = _f64(x)
    y = np.zeros_like(x)
    lib().orc_compcol_times(C.c_int32(x.shape[0]), _p(colptr), _p(rowind), _p(val), _p(x), _p(y))
    return y


def cg(colptr, rowind, val, b, x0=None, precond=1, max_iter=200, tol=1e-5):
    """IMLSolver defaults: tol 1e-5, maxite 200 (src/core/iml/imlsolver.C:72-75)."""
    b = _f64(b)
    x = np.zeros_like(b) if x0 is None else _f64(x0).copy()
    it, res = C.c_int(0), C.c_double(0.0)
    flag = lib().orc_cg(C.c_int32(b.shape[0]), _p(colptr), _p(rowind), _p(val), _p(b), _p(x), C.c_int(precond),
                        C.c_int(max_iter), C.c_double(tol), C.byref(it), C.byref(res))
    return x, flag, it.value, res.value


# --------------------------------------------------------------------------------------
# Engineering-model restatements (numpy on top of the C kernels)
# --------------------------------------------------------------------------------------

class Model:
    """Numbering and element data of a Problem (oofem_b200.inputfile.Problem)."""

    def __init__(self, pb):
        self.pb = pb
        self.etype = ETYPE[pb.elem_type]
        nn = pb.coords.shape[0]
        fixed = pb.fixed_mask()
        free = ~fixed.reshape(-1)
        # EngngModel::forceEquationNumbering (src/core/engngm.C:483-486)
        eq = np.zeros(nn * 3, dtype=np.int32)
        eq[free] = np.arange(1, int(free.sum()) + 1, dtype=np.int32)
        self.nodeeq = eq.reshape(nn, 3)
        self.neq = int(free.sum())
        self.loc = self.nodeeq[pb.conn - 1].reshape(pb.conn.shape[0], -1).astype(np.int32)
        self.matparams = pb.matparams()
        self.has_mises = any(m.kind == "misesmat" for m in pb.materials)
        self.state = mises_state(pb.conn.shape[0], self.etype) if self.has_mises else None
        self.colptr, self.rowind = compcol_build(self.loc, self.neq)

    def full_u(self, x, t):
        """Nodal displacement table [nnode,3]: free dofs from x, prescribed from the BCs at t."""
        u = self.pb.prescribed_values(t)
        m = self.nodeeq > 0
        u[m] = x[self.nodeeq[m] - 1]
        return u

    def stiffness(self):
        Ke = batch_stiffness(self.etype, self.pb.conn, self.pb.coords, self.pb.elem_mat, self.matparams, self.state)
        return compcol_assemble(self.loc, Ke, self.colptr, self.rowind)

    def internal_forces(self, u_nodes):
        ue = u_nodes[self.pb.conn - 1].reshape(self.pb.conn.shape[0], -1)
        fe = batch_internal_forces(self.etype, self.pb.conn, self.pb.coords, self.pb.elem_mat, self.matparams, ue,
                                   self.state)
        return assemble_vector(self.loc, fe, self.neq)

    def external_forces(self, t):
        f = self.pb.nodal_load_vector(t)
        out = np.zeros(self.neq)
        m = self.nodeeq > 0
        np.add.at(out, self.nodeeq[m] - 1, f[m])
        return out


def solve_linear_static(pb, precond=None, tol=None, max_iter=None):
    """LinearStatic::solveYourselfAt (src/sm/EngineeringModels/linearstatic.C:177-253), step 1:
    K from TangentAssembler, rhs = external - internal(prescribed), IML CG."""
    md = Model(pb)
    t = 1.0
    val = md.stiffness()
    rhs = md.external_forces(t) - md.internal_forces(md.full_u(np.zeros(md.neq), t))
    x, flag, it, res = cg(md.colptr, md.rowind, val, rhs,
                          precond=pb.params.get("lsprecond", 0) if precond is None else precond,
                          max_iter=pb.params.get("lsiter", 200) if max_iter is None else max_iter,
                          tol=pb.params.get("lstol", 1e-5) if tol is None else tol)
    return dict(model=md, val=val, rhs=rhs, x=x, u=md.full_u(x, t), flag=flag, iters=it, resid=res)


def solve_nonlinear_static(pb, rtol=1e-10, max_newton=50, lin_tol=1e-13, lin_iter=20000):
    """Load-controlled Newton-Raphson with full tangent re-assembly every iteration, the
    semantic content of NRSolver::solve (src/core/nrsolver.C) driven by StaticStructural /
    NonLinearStatic: per step t = 1..nsteps, iterate K du = f_ext(t) - f_int(u) with stresses
    always recomputed from the last committed state (total-strain update, misesmat.C:161-176),
    then commit (MaterialStatus::updateYourself).  Returns displacement tables per step."""
    md = Model(pb)
    x = np.zeros(md.neq)
    hist, iters = [], []
    for step in range(1, pb.params.get("nsteps", 1) + 1):
        t = float(step)
        fext = md.external_forces(t)
        # NRSolver::solve fetches the matrix BEFORE the first residual of the step
        # ("old tangent", src/core/nrsolver.C:252-259): elastic when nothing is yielding yet.
        val = md.stiffness()
        # IG_Tangent initial guess (src/sm/EngineeringModels/staticstructural.C:247-297):
        # K_old dx = -(sum_e Ke * increment of the prescribed dofs), then x += dx.
        du_p = pb.prescribed_values(t) - pb.prescribed_values(t - 1.0)
        if np.abs(du_p).max() > 0.0:
            Ke = batch_stiffness(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, md.state)
            due = du_p[pb.conn - 1].reshape(pb.conn.shape[0], -1)
            fex = assemble_vector(md.loc, np.einsum("eij,ej->ei", Ke, due), md.neq)
            dx, flag, n, res = cg(md.colptr, md.rowind, val, -fex, precond=1, max_iter=lin_iter, tol=lin_tol)
            x = x + dx
        dofid = np.nonzero(md.nodeeq > 0)[1]
        for it in range(max_newton):
            u_nodes = md.full_u(x, t)
            ue = u_nodes[pb.conn - 1].reshape(pb.conn.shape[0], -1)
            fe = batch_internal_forces(md.etype, pb.conn, pb.coords, pb.elem_mat, md.matparams, ue, md.state)
            fint = assemble_vector(md.loc, fe, md.neq)
            r = fext - fint
            # NRSolver::checkConvergence (src/core/nrsolver.C:752-760): per dof id,
            # sqrt( sum rhs^2 / ( sum RT^2 + element-by-element norm of the internal forces ) )
            ebe = (fe.reshape(fe.shape[0], -1, 3) ** 2).sum(axis=(0, 1))
            err = 0.0
            for g in range(3):
                m = dofid == g
                den = float(np.sum(fext[m] ** 2)) + float(ebe[g])
                num = float(np.sum(r[m] ** 2))
                err = max(err, np.sqrt(num / den) if den >= 1e-6 else np.sqrt(num))
            if err <= rtol and it > 0:
                break
            if it > 0:
                val = md.stiffness()          # manrmsteps 1: refresh every iteration (nrsolver.C:293-298)
            dx, flag, n, res = cg(md.colptr, md.rowind, val, r, precond=1, max_iter=lin_iter, tol=lin_tol)
            x = x + dx
        iters.append(it)
        if md.state is not None:
            md.internal_forces(md.full_u(x, t))
            mises_commit(md.state)
        hist.append(md.full_u(x, t))
    return dict(model=md, u_steps=hist, newton_iters=iters, x=x)
