"""Profiling driver: builds the bench mesh once and runs a few assemblies + SpMVs (for ncu -k ...)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
from oofem_b200.linsolver import CudaCG
import bench

nx = int(os.environ.get("NX", "250"))
ctx = capi.Context(0)
pb = bench.slab_problem(nx, 64, 64, 0, 1)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(a, device=dev)
nelem, neq = pb["conn"].shape[0], pb["neq"]
mp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
S = ElementSet(ctx, "lspace", t(pb["coords"]), t(pb["conn"]), t(np.zeros(nelem, np.int32)), mp, t(pb["loc"]), neq)
A = CudaCSR(ctx); A.buildInternalStructure(t(pb["loc"]), neq); S.bind(A)
for _ in range(3):
    A.zero(); S.assembleStiffness(A)
x = torch.ones(neq, dtype=torch.float64, device=dev); y = torch.zeros_like(x)
torch.cuda.synchronize()
for _ in range(4):
    A.times(x, y)
s = CudaCG(ctx).initializeFrom(dict(lstol=0.0, lsiter=6, lsprecond=1))
b = torch.ones(neq, dtype=torch.float64, device=dev); x0 = torch.zeros_like(b)
torch.cuda.synchronize()
s.solve(A, b, x0)
ctx.sync()
print("done", ctx.launches)
