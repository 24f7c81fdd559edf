"""Scratch: a few launches of the cluster assembly kernel at the full size (for ncu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oofem_b200 import capi
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (250, 64, 64))]
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
pb = bench.slab_problem(nx, ny, nz, 0, 1)
nelem, neq = pb["conn"].shape[0], pb["neq"]
matparams = np.array([[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
t = lambda a: torch.as_tensor(a, device=dev)
d = [t(pb["coords"]), t(pb["conn"]), t(np.zeros(nelem, np.int32)), t(pb["loc"])]
A = CudaCSR(ctx)
A.buildInternalStructure(d[3], neq)
S = ElementSet(ctx, "lspace", d[0], d[1], d[2], matparams, d[3], neq)
for _ in range(3):
    A.zero(); S.assembleStiffness(A)
ctx.sync()
