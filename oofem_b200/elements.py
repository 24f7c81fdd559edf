"""Batched element-evaluation hook: a homogeneous set of LSpace / LTRSpace elements kept
resident in HBM.  Replaces the per-element loop of EngngModel::assemble / assembleVector
(src/core/engngm.C:889-929) over StructuralElement::computeStiffnessMatrix and
giveInternalForcesVector (src/sm/Elements/structuralelement.C:575-643, 724-802)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, on_device, ptr

ETYPE = {"lspace": capi.LSPACE, "ltrspace": capi.LTRSPACE}
NEN = {capi.LSPACE: 8, capi.LTRSPACE: 4}
NGP = {capi.LSPACE: 8, capi.LTRSPACE: 1}


class ElementSet:
    def __init__(self, ctx: capi.Context, etype, coords, conn, matid, matparams, loc, neq: int, nodeeq=None):
        """loc [nelem, 3*nen]: the elements' location arrays (Element::giveLocationArray); or loc=None and
        nodeeq [nnode, 3]: the equation numbers of the dofs D_u, D_v, D_w of every node, from which the library forms
        the location arrays on the device (ob200_elemset_create_nodal)."""
        self.ctx = ctx
        self.etype = ETYPE[etype] if isinstance(etype, str) else int(etype)
        dev = on_device(coords)
        matparams = np.ascontiguousarray(matparams, dtype=np.float64).reshape(-1, capi.MATPARAM_STRIDE)
        if not dev:
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            conn = np.ascontiguousarray(conn, dtype=np.int32)
            matid = np.ascontiguousarray(matid, dtype=np.int32)
            loc = np.ascontiguousarray(loc, dtype=np.int32) if loc is not None else None
            nodeeq = np.ascontiguousarray(nodeeq, dtype=np.int32) if nodeeq is not None else None
            matparams_arg = matparams
        else:
            import torch
            matparams_arg = torch.as_tensor(matparams, device=coords.device)
        self.nnode, self.nelem = int(coords.shape[0]), int(conn.shape[0])
        self.nen, self.ngp = NEN[self.etype], NGP[self.etype]
        self.nd = 3 * self.nen
        self.neq = int(neq)
        if (loc is None) == (nodeeq is None):
            raise capi.OofemB200Error(capi.EINVAL, "give either the location arrays or the nodal equation numbers")
        if self.nelem and (int(conn.shape[1]) != self.nen or (loc is not None and int(loc.shape[1]) != self.nd)):
            raise capi.OofemB200Error(capi.EINVAL, "connectivity / location array shape does not match the element type")
        if nodeeq is not None and tuple(nodeeq.shape) != (self.nnode, 3):
            raise capi.OofemB200Error(capi.EINVAL, "nodal equation numbers must be [nnode, 3]")
        self.h = C.c_void_p()
        if loc is not None:
            check(lib().ob200_elemset_create(ctx.h, self.etype, self.nnode, ptr(coords), self.nelem, ptr(conn), ptr(matid),
                                             matparams.shape[0], ptr(matparams_arg), ptr(loc), self.neq, dev, C.byref(self.h)))
        else:
            check(lib().ob200_elemset_create_nodal(ctx.h, self.etype, self.nnode, ptr(coords), self.nelem, ptr(conn), ptr(matid),
                                                   matparams.shape[0], ptr(matparams_arg), ptr(nodeeq), self.neq, dev, C.byref(self.h)))

    @staticmethod
    def _out(like, shape):
        if like is None or isinstance(like, np.ndarray):
            return np.zeros(shape)
        import torch
        return torch.zeros(shape, dtype=torch.float64, device=like.device)

    def computeStiffnessMatrix(self, out=None):
        """Ke [nelem, nd, nd] for every element (TangentStiffness)."""
        if out is None:
            out = np.zeros((self.nelem, self.nd, self.nd))
        check(lib().ob200_elemset_stiffness(self.h, ptr(out), on_device(out)))
        return out

    def giveInternalForcesVector(self, u, want_gp=False):
        """fe [nelem, nd] for nodal displacements u [nnode, 3] (+ per-GP strain/stress)."""
        if isinstance(u, np.ndarray):
            u = np.ascontiguousarray(u, dtype=np.float64)
        fe = self._out(u, (self.nelem, self.nd))
        eps = self._out(u, (self.nelem * self.ngp, 6)) if want_gp else None
        sig = self._out(u, (self.nelem * self.ngp, 6)) if want_gp else None
        check(lib().ob200_elemset_internal_forces(self.h, ptr(u), ptr(fe), ptr(eps), ptr(sig), on_device(u)))
        return (fe, eps, sig) if want_gp else fe

    def bind(self, A):
        check(lib().ob200_elemset_bind(self.h, A.h))

    def assembleStiffness(self, A):
        """A += sum_e Ke (EngngModel::assemble with TangentAssembler), fused on the device."""
        check(lib().ob200_elemset_assemble_stiffness(self.h, A.h))

    def assembleInternalForces(self, u, f, eNorms=None):
        """f[neq] += sum_e fe (EngngModel::assembleVector with InternalForceAssembler).
        eNorms: optional numpy float64[3] receiving the element-by-element squared norms per dof id."""
        if isinstance(u, np.ndarray):
            u = np.ascontiguousarray(u, dtype=np.float64)
        check(lib().ob200_elemset_assemble_internal_forces(self.h, ptr(u), ptr(f), ptr(eNorms), on_device(u)))
        return f

    def assembleExtrapolatedForces(self, du, f):
        """f[neq] += sum_e Ke du_e (StaticStructural::assembleExtrapolatedForces)."""
        if isinstance(du, np.ndarray):
            du = np.ascontiguousarray(du, dtype=np.float64)
        check(lib().ob200_elemset_assemble_extrapolated_forces(self.h, ptr(du), ptr(f), on_device(du)))
        return f

    def updateYourself(self):
        """MaterialStatus::updateYourself for all Gauss points (temp -> committed)."""
        check(lib().ob200_elemset_commit(self.h))

    def state(self):
        st = np.zeros((self.nelem * self.ngp, capi.MISES_STATE_DOUBLES))
        check(lib().ob200_elemset_get_state(self.h, ptr(st), 0))
        return st

    def setState(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        check(lib().ob200_elemset_set_state(self.h, ptr(st), 0))

    def close(self):
        if self.h:
            lib().ob200_elemset_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
