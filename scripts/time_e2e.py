"""Scratch: where the end-to-end (host-pointer) step of bench.py spends its time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi
from oofem_b200.elements import ElementSet
from oofem_b200.linsolver import CudaCG
from oofem_b200.sparsemtrx import CudaCSR
import bench
ctx = capi.Context(0)
pb = bench.slab_problem(250, 64, 64, 0, 1)
nelem, neq = pb["conn"].shape[0], pb["neq"]
mp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hc, hn, hl, hm = pin(pb["coords"]), pin(pb["conn"]), pin(pb["loc"]), pin(np.zeros(nelem, np.int32))
hb, hx = pin(np.ones(neq)), pin(np.zeros(neq))
A = CudaCSR(ctx); A.buildInternalStructure(hl, neq)
solver = CudaCG(ctx, None).initializeFrom(dict(lstol=0.0, lsiter=50, lsprecond=1))
hq = pin(pb["nodeeq"].astype(np.int32))
nodal = "--nodal" in sys.argv          # upload nodal equation numbers instead of location arrays (ob200_elemset_create_nodal)
for rep in range(4):
    t = [time.perf_counter()]
    S = ElementSet(ctx, "lspace", hc, hn, hm, mp, None if nodal else hl, neq, nodeeq=hq if nodal else None); ctx.sync(); t.append(time.perf_counter())
    S.bind(A); ctx.sync(); t.append(time.perf_counter())
    A.zero(); S.assembleStiffness(A); ctx.sync(); t.append(time.perf_counter())
    hx[:] = 0; t.append(time.perf_counter())
    solver.solve(A, hb, hx); t.append(time.perf_counter())
    S.close(); ctx.sync(); t.append(time.perf_counter())
    names = ["create(h2d+incidence)", "bind(block schedule)", "assemble", "host x=0", "solve 50 it (h2d b,x; d2h x)", "close"]
    print("rep", rep, "  ".join(f"{n} {1e3 * (b - a):.2f} ms" for n, a, b in zip(names, t, t[1:])))
