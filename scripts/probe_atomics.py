"""Scratch probe: cost of the scatter phase alone (RED.F64 vs plain stores) on the bench mesh."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi
from oofem_b200.capi import check, lib
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
import bench
import ctypes as C

ctx = capi.Context(0)
pb = bench.slab_problem(250, 64, 64, 0, 1)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(a, device=dev)
nelem, neq = pb["conn"].shape[0], pb["neq"]
mp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
S = ElementSet(ctx, "lspace", t(pb["coords"]), t(pb["conn"]), t(np.zeros(nelem, np.int32)), mp, t(pb["loc"]), neq)
A = CudaCSR(ctx); A.buildInternalStructure(t(pb["loc"]), neq); S.bind(A)
f = lib().ob200_debug_probe_scatter
f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
for mode, name in ((2, "slot read only"), (1, "plain store"), (0, "RED.F64")):
    for bps in (4, 8):
        for rep in range(3):
            A.zero(); ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext); check(f(S.h, A.h, mode, bps)); e1.record(ext); ctx.sync()
        print(f"{name:16s} blocks/SM={bps} {e0.elapsed_time(e1):8.3f} ms")
