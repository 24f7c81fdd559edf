"""Host-side mirror of the reference's SparseMtrx interface (src/core/sparsemtrx.h) for
the new sparse matrix type "cudacsr" -- same method names, argument meaning and error
behaviour as CompCol (src/core/compcol.C), storage and arithmetic on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, on_device, ptr


class CudaCSR:
    """SparseMtrx "cudacsr" (SMT_CudaCSR).  Indices passed in follow OOFEM: location arrays
    are 1-based with 0 meaning a prescribed dof; at(i, j) is 1-based."""

    def __init__(self, ctx: capi.Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        check(lib().ob200_csr_create(ctx.h, C.byref(self.h)))

    # -- SparseMtrx::buildInternalStructure(EngngModel*, int di, const UnknownNumberingScheme&)
    #    (compcol.C:167-260).  The engineering model is represented by what the method reads
    #    from it: the element location arrays and the number of equations.
    def buildInternalStructure(self, loc, neq: int) -> int:
        if isinstance(loc, np.ndarray):
            loc = np.ascontiguousarray(loc, dtype=np.int32)
        nelem, nd = (int(loc.shape[0]), int(loc.shape[1])) if loc.ndim == 2 else (0, 1)
        check(lib().ob200_csr_build_structure(self.h, neq, nelem, nd, ptr(loc), on_device(loc)))
        return True

    # -- SparseMtrx::assemble(const IntArray &loc, const FloatMatrix &mat) (compcol.C:263-299);
    #    also accepts a batch: loc [nelem][nd], mat [nelem][nd][nd]
    def assemble(self, loc, mat) -> int:
        if isinstance(loc, np.ndarray):
            loc = np.ascontiguousarray(loc, dtype=np.int32)
            mat = np.ascontiguousarray(mat, dtype=np.float64)
        if loc.ndim == 1:
            nelem, nd = 1, int(loc.shape[0])
        else:
            nelem, nd = int(loc.shape[0]), int(loc.shape[1])
        if int(np.prod(mat.shape)) != nelem * nd * nd:
            raise capi.OofemB200Error(capi.EINVAL, "dimension of 'k' and 'loc' mismatch")   # compcol.C:268-270
        check(lib().ob200_csr_assemble(self.h, nelem, nd, ptr(loc), ptr(mat), on_device(loc)))
        return 1

    # -- SparseMtrx::times(const FloatArray &x, FloatArray &answer) (compcol.C:119-134)
    def times(self, x, answer=None):
        n = self.giveNumberOfRows()
        if int(np.prod(x.shape)) != n:
            raise capi.OofemB200Error(capi.EINVAL, "incompatible dimensions")               # compcol.C:121-123
        if answer is None:
            if isinstance(x, np.ndarray):
                answer = np.zeros(n)
            else:
                import torch
                answer = torch.zeros(n, dtype=torch.float64, device=x.device)
        check(lib().ob200_csr_times(self.h, ptr(x), ptr(answer), on_device(x)))
        return answer

    def timesT(self, x, answer=None):                       # SparseMtrx::timesT (compcol.C:146-163)
        n = self.giveNumberOfRows()
        if int(np.prod(x.shape)) != n:
            raise capi.OofemB200Error(capi.EINVAL, "incompatible dimensions")
        if answer is None:
            if isinstance(x, np.ndarray):
                answer = np.zeros(n)
            else:
                import torch
                answer = torch.zeros(n, dtype=torch.float64, device=x.device)
        check(lib().ob200_csr_times_t(self.h, ptr(x), ptr(answer), on_device(x)))
        return answer

    def times_scalar(self, s: float):                       # SparseMtrx::times(double) (compcol.C:159-164)
        check(lib().ob200_csr_scale(self.h, float(s)))

    def zero(self):                                         # compcol.C:339-344
        check(lib().ob200_csr_zero(self.h))

    def at(self, i: int, j: int) -> float:                  # compcol.C:376-390 (1-based)
        v = C.c_double(0.0)
        check(lib().ob200_csr_at(self.h, i, j, C.byref(v)))
        return v.value

    def isAllocatedAt(self, i: int, j: int) -> bool:         # SparseMtrx::isAllocatedAt: (i,j) is in the sparse structure
        v = C.c_double(0.0)
        return lib().ob200_csr_at(self.h, i, j, C.byref(v)) == 0

    def giveNumberOfRows(self) -> int:
        return lib().ob200_csr_rows(self.h)

    giveNumberOfColumns = giveNumberOfRows

    def giveNumberOfNonzeros(self) -> int:
        return lib().ob200_csr_nnz(self.h)

    def giveVersion(self) -> int:
        return lib().ob200_csr_version(self.h)

    def isAsymmetric(self) -> bool:
        return True

    def giveClassName(self) -> str:
        return "CudaCSR"

    # -- inspection helpers (tests, export)
    def structure(self):
        """(rowptr[neq+1], colind[nnz]) as numpy int32 -- the integers of CompCol's colptr/rowind."""
        rp = np.zeros(self.giveNumberOfRows() + 1, dtype=np.int32)
        ci = np.zeros(max(self.giveNumberOfNonzeros(), 1), dtype=np.int32)
        check(lib().ob200_csr_get_structure(self.h, ptr(rp), ptr(ci), 0))
        return rp, ci[:self.giveNumberOfNonzeros()]

    def values(self, device: bool = False):
        """val[nnz]: numpy array, or (device=True) a torch tensor on the context's GPU (device-to-device copy)."""
        n = self.giveNumberOfNonzeros()
        if device:
            import torch
            v = torch.zeros(max(n, 1), dtype=torch.float64, device=torch.device("cuda", self.ctx.device))
            check(lib().ob200_csr_get_values(self.h, ptr(v), 1))
            return v[:n]
        v = np.zeros(max(n, 1))
        check(lib().ob200_csr_get_values(self.h, ptr(v), 0))
        return v[:n]

    def set_values(self, v):
        if isinstance(v, np.ndarray):
            v = np.ascontiguousarray(v, dtype=np.float64)
        check(lib().ob200_csr_set_values(self.h, ptr(v), on_device(v)))

    def close(self):
        if self.h:
            lib().ob200_csr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
