"""Scratch: time the assembly kernels at the full size, compare the cluster kernel against the round-1 gather kernel."""
import os, sys, time, json
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
vals = {}
for path in (("cluster", "gather") if not os.environ.get("OB200_ONLY") else (os.environ["OB200_ONLY"],)):
    if path == "gather":
        os.environ["OB200_ASSEMBLY"] = "gather"
    else:
        os.environ.pop("OB200_ASSEMBLY", None)
    S = ElementSet(ctx, "lspace", d[0], d[1], d[2], matparams, d[3], neq)
    t0 = time.time(); S.bind(A); ctx.sync(); tb = time.time() - t0
    for _ in range(3):
        A.zero(); S.assembleStiffness(A)
    ctx.sync()
    ctx.profile_reset(); ctx.set_profiling(True)
    for _ in range(5):
        A.zero(); S.assembleStiffness(A)
    ctx.sync(); ctx.set_profiling(False)
    prof = ctx.profile_report()
    vals[path] = A.values().copy() if isinstance(A.values(), np.ndarray) else A.values()
    print(path, "bind_s %.4f" % tb, {k: (round(v[0] / v[1], 4), v[1]) for k, v in prof.items()})
    S.close()
if len(vals) == 2:
  a, b = np.asarray(vals["cluster"]), np.asarray(vals["gather"])
  print("relerr cluster vs gather:", float(np.abs(a - b).max() / np.abs(b).max()), "nnz", a.size)
