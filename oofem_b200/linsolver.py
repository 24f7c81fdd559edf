"""Host-side mirror of SparseLinearSystemNM (src/core/sparselinsystemnm.h) for the new
linear solver "cudacg": the IML++ CG of IMLSolver (src/core/iml/imlsolver.C) on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, on_device, ptr

# ConvergedReason (src/core/convergedreason.h)
CR_CONVERGED = 0
CR_DIVERGED_ITS = 1


class CudaCG:
    """SparseLinearSystemNM "cudacg" (ST_CudaCG).  Input keywords are IMLSolver's
    (imlsolver.h:48-52): lstol (default 1e-5), lsiter (200), lsprecond (0 void, 1 diagonal)."""

    def __init__(self, ctx: capi.Context, comm=None):
        self.ctx = ctx
        self.comm = comm
        self.tol = 1.0e-5
        self.maxite = 200
        self.precondType = capi.PRECOND_VOID
        self.last_iterations = 0
        self.last_residual = 0.0

    def initializeFrom(self, ir: dict):
        # IMLSolver::initializeFrom (imlsolver.C:66-98)
        stype = int(ir.get("stype", 0))
        if stype != 0:
            raise capi.OofemB200Error(capi.EINVAL, "unknown lsover type")          # only IML_ST_CG here
        self.tol = float(ir.get("lstol", 1.0e-5))
        self.maxite = int(ir.get("lsiter", 200))
        self.precondType = int(ir.get("lsprecond", 0))
        if self.precondType not in (capi.PRECOND_VOID, capi.PRECOND_DIAG):
            raise capi.OofemB200Error(capi.EINVAL, "unknown preconditioner type")  # imlsolver.C:93
        return self

    def solve(self, A, b, x):
        """IMLSolver::solve(SparseMtrx &A, FloatArray &b, FloatArray &x) (imlsolver.C:101-146):
        x holds the initial guess and receives the solution.  Returns ConvergedReason."""
        n = A.giveNumberOfRows()
        if int(np.prod(x.shape)) != int(np.prod(b.shape)) or int(np.prod(b.shape)) != n:
            raise capi.OofemB200Error(capi.EINVAL, "size mismatch")                 # imlsolver.C:105-107
        if isinstance(b, np.ndarray):
            b = np.ascontiguousarray(b, dtype=np.float64)
            if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
                raise capi.OofemB200Error(capi.EINVAL, "x must be a contiguous float64 array (it is updated in place)")
        it, res = C.c_int(0), C.c_double(0.0)
        if self.comm is None:
            flag = check(lib().ob200_cg_solve(A.h, ptr(b), ptr(x), self.precondType, self.maxite, self.tol,
                                              C.byref(it), C.byref(res), on_device(b)))
        else:
            flag = check(lib().ob200_cg_solve_dist(A.h, self.comm.h, ptr(b), ptr(x), self.precondType, self.maxite,
                                                   self.tol, C.byref(it), C.byref(res), on_device(b)))
        self.last_iterations, self.last_residual = it.value, res.value
        return CR_CONVERGED if flag == 0 else CR_DIVERGED_ITS

    def giveClassName(self) -> str:
        return "CudaCG"
