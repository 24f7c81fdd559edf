#!/usr/bin/env python3
"""The BASELINE.json configurations that are not the bench line, at full size on one B200:
config 2 (4M-element LTRSpace linear elastic) and config 3 (1M-hex MisesMat NonLinearStatic with
per-iteration tangent reassembly).  Prints one JSON line per configuration with timings (CUDA events
per kernel through the context profiler) and size-independent property checks.

    python scripts/run_configs.py [ltrspace] [mises]
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oofem_b200 import capi, meshgen
from oofem_b200.elements import ElementSet
from oofem_b200.engng import StaticStructural
from oofem_b200.inputfile import DirichletBC, Material, Problem
from oofem_b200.linsolver import CudaCG
from oofem_b200.sparsemtrx import CudaCSR

relerr = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def timed(ctx, fn):
    ctx.sync(); t0 = time.perf_counter(); r = fn(); ctx.sync()
    return r, (time.perf_counter() - t0) * 1e3


def ltrspace(ctx):
    nx, ny, nz = 110, 78, 78                      # 6 * 110*78*78 = 4,015,440 tetrahedra
    coords, conn = meshgen.tet_beam(nx, ny, nz)
    coords = meshgen.perturb(coords, 0.1 / ny, seed=5)
    fixed, _ = meshgen.cantilever_bcs(coords, float(nx) / ny)
    mask = np.zeros((coords.shape[0], 3), bool); mask[fixed - 1] = True
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(a, device=dev)
    S = ElementSet(ctx, "ltrspace", t(coords), t(conn), t(np.zeros(conn.shape[0], np.int32)), [[1, 210e3, 0.3, 0, 0, 0, 0, 0]], t(loc), neq)
    A = CudaCSR(ctx)
    _, t_struct = timed(ctx, lambda: A.buildInternalStructure(t(loc), neq))
    S.bind(A)
    A.zero(); S.assembleStiffness(A)              # warm-up
    ts = []
    ctx.profile_reset(); ctx.set_profiling(True)
    for _ in range(5):
        _, ms = timed(ctx, lambda: (A.zero(), S.assembleStiffness(A)))
        ts.append(ms)
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    v1 = A.values(device=True).clone()
    A.zero(); S.assembleStiffness(A); ctx.sync()
    reproducible = bool(torch.equal(v1, A.values(device=True)))
    del v1
    nnz = A.giveNumberOfNonzeros()
    rng = np.random.default_rng(1)
    x, y = t(rng.normal(size=neq)), t(rng.normal(size=neq))
    Ax, Ay = A.times(x), A.times(y)
    ctx.sync()
    sym = abs(float(x @ Ay - y @ Ax)) / abs(float(x @ Ay))
    f = t(np.zeros(neq))
    S.assembleInternalForces(t(np.ones_like(coords)), f); ctx.sync()
    rigid = float(f.abs().max()) / float(torch.as_tensor(A.values()).abs().max())
    b = t(rng.normal(size=neq)); xs = torch.zeros(neq, dtype=torch.float64, device=dev)
    solver = CudaCG(ctx).initializeFrom(dict(lstol=0.0, lsiter=200, lsprecond=1))
    solver.solve(A, b, xs)
    xs.zero_()
    _, t_cg = timed(ctx, lambda: solver.solve(A, b, xs))
    s2 = CudaCG(ctx).initializeFrom(dict(lstol=1e-9, lsiter=50000, lsprecond=1))
    xs.zero_()
    torch.cuda.synchronize()
    flag = s2.solve(A, b, xs); ctx.sync()
    Axs = A.times(xs); ctx.sync()         # the product runs on the context's stream, torch on its own
    res = float(torch.linalg.norm(Axs - b) / torch.linalg.norm(b))
    return {"config": "BASELINE configs[2] (one partition): LTRSpace linear elastic, structured Kuhn split, perturbed nodes",
            "nelem": int(conn.shape[0]), "neq": int(neq), "nnz": int(nnz), "structure_build_ms": round(t_struct, 2),
            "assembly_ms": round(min(ts), 3), "elements_per_s": conn.shape[0] / (min(ts) * 1e-3),
            "assembly_kernels_ms_avg": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}, "bit_reproducible": reproducible,
            "pcg_iters_per_s": 200 / (t_cg * 1e-3), "checks": {"symmetry_rel": sym, "rigid_translation_force_rel": rigid,
            "cg_converged": flag == 0, "cg_iters": s2.last_iterations, "true_residual_rel": res}}


def mises(ctx):
    nx, ny, nz = [int(v) for v in os.environ.get("MISES_DIMS", "250,64,64").split(",")]
    coords, conn = meshgen.hex_beam(nx, ny, nz)
    lx = float(nx) / ny
    left = np.nonzero(coords[:, 0] < 1e-9)[0] + 1
    right = np.nonzero(coords[:, 0] > lx - 1e-9)[0] + 1
    pb = Problem(engng="staticstructural", params=dict(nsteps=2, rtolf=1e-6, maxiter=30, lstol=1e-9, lsiter=50000, lsprecond=1),
                 coords=coords, elem_type="lspace", conn=conn, elem_mat=np.zeros(conn.shape[0], np.int32),
                 materials=[Material("misesmat", 210e3, 0.3, sig0=250.0, H=2100.0)])
    pb.ltfs[1] = ("const", 1.0)
    pb.ltfs[2] = ("pwl", [0.0, 1.0, 2.0], [0.0, 1.0, 1.25])
    pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, left))
    # stretch the bar: 0.1 % strain in step 1 (yield strain 0.119 %: plastic only near the clamped ends),
    # 0.125 % in step 2 (the whole bar yields)
    pb.bcs.append(DirichletBC([1], [1.0e-3 * lx], 2, right))
    ctx.profile_reset(); ctx.set_profiling(True)
    t0 = time.perf_counter()
    em = StaticStructural(ctx, pb)
    steps = []
    for step in (1, 2):
        ts = time.perf_counter()
        try:
            u = em.solveYourselfAt(step)
        except capi.OofemB200Error as e:
            print("NR failed:", e, "trace:", em.trace, file=sys.stderr)
            raise
        steps.append({"step": step, "newton_iterations": em.iterations[-1], "wall_s": round(time.perf_counter() - ts, 2),
                      "cg_iterations_last_solve": em.linSolver.last_iterations, "u_max": float(np.abs(u).max())})
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    st = em.domain.elems.state()
    kappa = st.reshape(-1, 29)[:, 6] if st.size else np.zeros(1)
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]
    return {"config": "BASELINE configs[3]: MisesMat J2 plasticity, StaticStructural/NRSolver, tangent reassembled every iteration, 1M hex",
            "nelem": int(conn.shape[0]), "neq": int(em.domain.neq), "steps": steps, "tangent_assemblies": em.tangent_assemblies,
            "newton_trace(step, iteration, force_error, cg_iters_prev_solve)": [(a, b, float(f"{c:.3e}"), d) for a, b, c, d in em.trace],
            "wall_s_total": round(time.perf_counter() - t0, 2),
            "kernels_ms_avg": {k: round(v[0] / max(v[1], 1), 4) for k, v in top}, "kernel_launches": {k: v[1] for k, v in top},
            "checks": {"plastic_gauss_points_frac": float(np.mean(kappa > 0)) if st.size else None,
                       "newton_converged": True}}


if __name__ == "__main__":
    ctx = capi.Context(0)
    want = sys.argv[1:] or ["ltrspace", "mises"]
    for name in want:
        print(json.dumps({"ltrspace": ltrspace, "mises": mises}[name](ctx)), flush=True)
