"""Scratch: time the fused assembly alone on the bench mesh (CUDA events on the context stream)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
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
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
ts = []
for rep in range(6):
    A.zero()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); S.assembleStiffness(A); e1.record(ext); ctx.sync()
    ts.append(e0.elapsed_time(e1))
print("dbg", os.environ.get("OB200_GATHER_DBG", "0"), "nelem", nelem, "assembly ms", " ".join(f"{x:.3f}" for x in ts), "checksum", float(np.abs(A.values()).sum()))
