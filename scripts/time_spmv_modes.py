"""Scratch: time the plain / fused-dot / fused-halo modes of the streamed SpMV on one GPU."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi
from oofem_b200.capi import ptr
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
import bench
ctx = capi.Context(0)
pb = bench.slab_problem(250, 64, 64, 0, 1)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(a, device=dev)
nelem, neq = pb["conn"].shape[0], pb["neq"]
mp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
S = ElementSet(ctx, "lspace", t(pb["coords"]), t(pb["conn"]), t(np.zeros(nelem, np.int32)), mp, t(pb["loc"]), neq)
A = CudaCSR(ctx); A.buildInternalStructure(t(pb["loc"]), neq); S.bind(A); A.zero(); S.assembleStiffness(A)
x = torch.rand(neq, dtype=torch.float64, device=dev); y = torch.zeros_like(x)
ms = (C.c_float * 3)()
L = C.CDLL(capi.LIB_PATH)
L.ob200_debug_spmv_modes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
for rep in range(2):
    rc = L.ob200_debug_spmv_modes(A.h, ptr(x), ptr(y), 50, ms)
    print("rc", rc, "spmv ms: plain %.4f fused-dot %.4f fused-halo(route=-1) %.4f" % tuple(ms))
