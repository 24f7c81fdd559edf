"""Scratch: time the internal-force assembly at 1M hex, IsoLE vs MisesMat (elastic / plastic state)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi, meshgen
from oofem_b200.elements import ElementSet
ctx = capi.Context(0)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(a, device=dev)
nx, ny, nz = 250, 64, 64
coords, conn = meshgen.hex_beam(nx, ny, nz)
mask = np.zeros((coords.shape[0], 3), bool); mask[:(ny + 1) * (nz + 1)] = True
nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
loc = meshgen.location_arrays(conn, nodeeq)
u = np.zeros_like(coords); u[:, 0] = 2.0e-3 * coords[:, 0]; u[:, 1] = -0.5e-3 * coords[:, 1]
for name, mp, scale in (("isole", [[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], 1.0),
                        ("mises elastic", [[capi.MAT_MISES, 210e3, 0.3, 250.0, 2100.0, 0.2, 30.0, 0]], 1e-3),
                        ("mises plastic", [[capi.MAT_MISES, 210e3, 0.3, 250.0, 2100.0, 0.2, 30.0, 0]], 1.0)):
    S = ElementSet(ctx, "lspace", t(coords), t(conn), t(np.zeros(conn.shape[0], np.int32)), np.array(mp, dtype=np.float64), t(loc), neq)
    f = t(np.zeros(neq)); ud = t(u * scale)
    S.assembleInternalForces(ud, f)
    ctx.profile_reset(); ctx.set_profiling(True)
    for _ in range(4):
        S.assembleInternalForces(ud, f)
    ctx.sync(); ctx.set_profiling(False)
    print(name, {k: round(v[0] / max(v[1], 1), 4) for k, v in ctx.profile_report().items()})
    S.close()
