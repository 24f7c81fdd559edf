"""Engineering-model drivers over the GPU path (the callers of the hot path).

LinearStatic follows src/sm/EngineeringModels/linearstatic.C:177-253, StaticStructural
follows src/sm/EngineeringModels/staticstructural.C:224-345 with NRSolver::solve
(src/core/nrsolver.C:215-345) in load control.  Both are thin host code: every loop over
elements, every matrix / vector operation of size neq runs on the device through the C ABI.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .elements import ElementSet
from .inputfile import Problem, SMT_CUDACSR, ST_CUDACG
from .linsolver import CR_CONVERGED, CudaCG
from .meshgen import equation_numbers, location_arrays
from .sparsemtrx import CudaCSR


class Domain:
    """Numbering + device-resident element data of a Problem (one homogeneous element set)."""

    def __init__(self, ctx: capi.Context, pb: Problem):
        self.ctx, self.pb = ctx, pb
        # EngngModel::forceEquationNumbering (src/core/engngm.C:483-486)
        self.nodeeq, self.neq = equation_numbers(pb.coords.shape[0], pb.fixed_mask())
        self.loc = location_arrays(pb.conn, self.nodeeq)
        self.free = self.nodeeq > 0
        self.elems = ElementSet(ctx, pb.elem_type, pb.coords, pb.conn, pb.elem_mat, pb.matparams(), self.loc, self.neq)

    def full_u(self, x: np.ndarray, t: float) -> np.ndarray:
        u = self.pb.prescribed_values(t)
        u[self.free] = x[self.nodeeq[self.free] - 1]
        return u

    def external_forces(self, t: float) -> np.ndarray:
        f = self.pb.nodal_load_vector(t)
        out = np.zeros(self.neq)
        np.add.at(out, self.nodeeq[self.free] - 1, f[self.free])
        return out


def _check_solver_selection(pb: Problem):
    """An unmodified .in selects the GPU path with `lstype cudacg smtype cudacsr` (or the
    numeric enum values); any other explicit selection is refused -- no CPU fallback."""
    ls, sm = pb.params.get("lstype", ST_CUDACG), pb.params.get("smtype", SMT_CUDACSR)
    if ls != ST_CUDACG or sm != SMT_CUDACSR:
        raise capi.OofemB200Error(capi.EINVAL, f"lstype {ls} / smtype {sm}: this package provides only "
                                               f"lstype {ST_CUDACG} (cudacg) with smtype {SMT_CUDACSR} (cudacsr)")


class LinearStatic:
    def __init__(self, ctx: capi.Context, pb: Problem, strict_selection: bool = False):
        if strict_selection:
            _check_solver_selection(pb)
        self.ctx, self.pb = ctx, pb
        self.domain = Domain(ctx, pb)
        self.stiffnessMatrix = None
        self.nMethod = CudaCG(ctx).initializeFrom(pb.params)
        self.displacementVector = None

    def solveYourselfAt(self, t: float = 1.0):
        d = self.domain
        if self.stiffnessMatrix is None:                        # initFlag branch, linearstatic.C:184-203
            self.stiffnessMatrix = CudaCSR(self.ctx)
            self.stiffnessMatrix.buildInternalStructure(d.loc, d.neq)
            d.elems.assembleStiffness(self.stiffnessMatrix)
        load = d.external_forces(t)                             # ExternalForceAssembler
        internal = np.zeros(d.neq)                              # InternalForceAssembler (Dirichlet b.c.)
        d.elems.assembleInternalForces(d.full_u(np.zeros(d.neq), t), internal)
        load -= internal
        x = np.zeros(d.neq)
        s = self.nMethod.solve(self.stiffnessMatrix, load, x)
        if s != CR_CONVERGED:
            raise capi.OofemB200Error(capi.EINVAL, "No success in solving system.")    # linearstatic.C:247-249
        self.displacementVector = x
        self.loadVector = load
        return d.full_u(x, t)


class StaticStructural:
    """Load-controlled Newton-Raphson; `manrmsteps 1` semantics (tangent re-assembled every
    iteration, nrsolverAccelNRM with MANRMSteps = 1, nrsolver.C:122-125, 293-298)."""

    def __init__(self, ctx: capi.Context, pb: Problem, lin_tol: float = 1e-12, lin_iter: int = 20000):
        self.ctx, self.pb = ctx, pb
        self.domain = Domain(ctx, pb)
        d = self.domain
        self.stiffnessMatrix = CudaCSR(ctx)
        self.stiffnessMatrix.buildInternalStructure(d.loc, d.neq)
        self.linSolver = CudaCG(ctx).initializeFrom(dict(lstol=pb.params.get("lstol", lin_tol),
                                                         lsiter=pb.params.get("lsiter", lin_iter), lsprecond=1))
        self.rtolf = pb.params.get("rtolf", pb.params.get("rtolv", 1e-3))     # nrsolver.C:137-147
        self.nsmax = pb.params.get("maxiter", 100)
        self.solution = np.zeros(d.neq)
        self.iterations = []
        self.trace = []                 # (step, iteration, force error, CG iterations of the previous solve)
        self.tangent_assemblies = 0

    def updateMatrix(self):
        self.stiffnessMatrix.zero()
        self.domain.elems.assembleStiffness(self.stiffnessMatrix)
        self.tangent_assemblies += 1

    def internal_forces(self, t):
        f = np.zeros(self.domain.neq)
        self.internalForcesEBENorm = np.zeros(3)
        self.domain.elems.assembleInternalForces(self.domain.full_u(self.solution, t), f, self.internalForcesEBENorm)
        return f

    def force_error(self, rhs, fext):
        """NRSolver::checkConvergence (nrsolver.C:594-760): per dof id group (D_u, D_v, D_w)
        forceErr = sqrt( sum rhs^2 / ( sum RT^2 + internalForcesEBENorm ) ); the largest governs."""
        d = self.domain
        dofid = np.nonzero(d.free)[1]              # dof id of every equation, in equation order
        err = 0.0
        for g in range(3):
            m = dofid == g
            num = float(np.sum(rhs[m] ** 2))
            den = float(np.sum(fext[m] ** 2)) + float(self.internalForcesEBENorm[g])
            err = max(err, np.sqrt(num / den) if den >= 1e-6 else np.sqrt(num))     # nrsolver_ERROR_NORM_SMALL_NUM
        return err

    def solveYourselfAt(self, step: int):
        d, pb = self.domain, self.pb
        t = float(step)
        # "old tangent" fetched before the residual (nrsolver.C:252-259)
        self.updateMatrix()
        # IG_Tangent initial guess for the increment of prescribed dofs (staticstructural.C:247-297)
        du_p = pb.prescribed_values(t) - pb.prescribed_values(t - 1.0)
        if np.abs(du_p).max() > 0.0:
            fex = np.zeros(d.neq)
            d.elems.assembleExtrapolatedForces(du_p, fex)
            inc = np.zeros(d.neq)
            self.linSolver.solve(self.stiffnessMatrix, -fex, inc)
            self.solution = self.solution + inc
        fext = d.external_forces(t)
        nite = 0
        while True:
            fint = self.internal_forces(t)
            rhs = fext - fint
            err = self.force_error(rhs, fext)
            self.trace.append((step, nite, err, self.linSolver.last_iterations))
            if err <= self.rtolf and ( nite > 0 or err == 0.0 ):
                break
            if nite >= self.nsmax:
                raise capi.OofemB200Error(capi.EINVAL, "Maximum number of iterations reached without convergence")
            if nite > 0:
                self.updateMatrix()
            ddx = np.zeros(d.neq)
            self.linSolver.solve(self.stiffnessMatrix, rhs, ddx)
            self.solution = self.solution + ddx
            nite += 1
        self.iterations.append(nite)
        d.elems.updateYourself()                               # EngngModel::updateYourself
        return d.full_u(self.solution, t)

    def solveYourself(self):
        out = []
        for step in range(1, self.pb.params.get("nsteps", 1) + 1):
            out.append(self.solveYourselfAt(step))
        return out


def solve(ctx: capi.Context, pb: Problem):
    """Run the problem the way `oofem -f file.in` would and return the nodal displacement
    table(s) [nnode, 3] per step."""
    if pb.engng == "linearstatic":
        return [LinearStatic(ctx, pb).solveYourselfAt(1.0)]
    has_mises = any(m.kind == "misesmat" for m in pb.materials)
    if not has_mises and pb.params.get("nsteps", 1) == 1:
        em = StaticStructural(ctx, pb)
        return em.solveYourself()
    return StaticStructural(ctx, pb).solveYourself()
