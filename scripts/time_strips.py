"""Scratch: time the general-tangent assembly (assemble_strips.cu) and the LTRSpace node-row assembly (assemble_tet.cu) at full size.
    python scripts/time_strips.py [mises|tet] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi, meshgen
from oofem_b200.elements import ElementSet
from oofem_b200.sparsemtrx import CudaCSR
what = sys.argv[1] if len(sys.argv) > 1 else "mises"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = capi.Context(0)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(a, device=dev)
if what in ("mises", "isole"):
    nx, ny, nz = 250, 64, 64
    coords, conn = meshgen.hex_beam(nx, ny, nz)
    mp = np.array([[capi.MAT_MISES, 210e3, 0.3, 250.0, 2100.0, 0.2, 30.0, 0]], dtype=np.float64)
    if what == "isole":          # the general-tangent path forced onto a linear elastic set
        mp = np.array([[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
        os.environ["OB200_ASSEMBLY"] = "strips"
    et = "lspace"
else:
    nx, ny, nz = 110, 78, 78
    coords, conn = meshgen.tet_beam(nx, ny, nz)
    coords = meshgen.perturb(coords, 0.1 / ny, seed=5)
    mp = np.array([[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    et = "ltrspace"
mask = np.zeros((coords.shape[0], 3), bool)
mask[:(ny + 1) * (nz + 1)] = True
nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
loc = meshgen.location_arrays(conn, nodeeq)
S = ElementSet(ctx, et, t(coords), t(conn), t(np.zeros(conn.shape[0], np.int32)), mp, t(loc), neq)
A = CudaCSR(ctx); A.buildInternalStructure(t(loc), neq); S.bind(A)
if what == "mises":      # a plastic increment (not committed): every Gauss point gets an algorithmic tangent
    u = np.zeros_like(coords); u[:, 0] = 2.0e-3 * coords[:, 0]; u[:, 1] = -0.5e-3 * coords[:, 1]
    f = t(np.zeros(neq)); S.assembleInternalForces(t(u), f)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
ctx.profile_reset(); ctx.set_profiling(True)
ts = []
for rep in range(reps):
    A.zero()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); S.assembleStiffness(A); e1.record(ext); ctx.sync()
    ts.append(e0.elapsed_time(e1))
ctx.set_profiling(False)
prof = ctx.profile_report()
print(what, "nelem", conn.shape[0], "assembly ms", " ".join(f"{x:.3f}" for x in ts),
      {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()})
