#!/usr/bin/env python3
"""End-to-end through OOFEM's own executable: `oofem_cuda -f big.in` (plugin/_build, cudacsr + cudacg + batched hooks) against
`oofem -f big.in` (oracle/_ref/oofem_omp, the unmodified reference, _OPENMP build on all host threads) on the same generated
LinearStatic input: a structured LSpace cantilever, IML CG with the diagonal preconditioner to lstol (both solve to
convergence, so the iteration counts are the solvers' own).  Prints one JSON object.

    python scripts/e2e_executable.py [nx ny nz] [--lstol 1e-3] [--skip-reference]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_case(td, dims, lstol):
    from oofem_b200 import meshgen
    from oofem_b200.inputfile import DirichletBC, Material, NodalLoad, Problem, write_input
    nx, ny, nz = dims
    h = 1.0 / ny
    coords, conn = meshgen.hex_beam(nx, ny, nz, nx * h, 1.0, nz * h)
    fixed, tip = meshgen.cantilever_bcs(coords, nx * h)
    out = {}
    for tag, (ls, sm) in (("ref", (1, 2)), ("cuda", (9, 11))):
        pb = Problem(title="executable e2e", outfile=os.path.join(td, tag + ".out"), engng="linearstatic",
                     params=dict(nsteps=1, lstype=ls, smtype=sm, lstol=lstol, lsiter=1000000, lsprecond=1), coords=coords,
                     elem_type="lspace", conn=conn, elem_mat=np.zeros(conn.shape[0], np.int32), materials=[Material("isole", 210e3, 0.3)])
        pb.ltfs[1] = ("const", 1.0)
        pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, fixed))
        pb.loads.append(NodalLoad([3], [-1.0], 1, tip))
        fn = os.path.join(td, tag + ".in")
        write_input(fn, pb)
        # node / element output of the first entity only: the .out of a million elements would time the disk
        txt = open(fn).read().replace("OutputManager tstep_all dofman_all element_all", "OutputManager tstep_all dofman_output {%d} element_output {1}" % int(tip[0]))
        open(fn, "w").write(txt)
        out[tag] = fn
    return out, conn.shape[0], int(tip[0])


def run(exe, fn, env=None):
    t0 = time.time()
    r = subprocess.run([exe, "-f", fn], capture_output=True, text=True, errors="replace", env=env, cwd=os.path.dirname(fn))
    wall = time.time() - t0
    log = r.stdout + r.stderr
    res = {"wall_s": round(wall, 3), "rc": r.returncode}
    m = re.search(r"nite (\d+), achieved tol\. ([-+.\de]+)", log)
    if m:
        res["cg_iterations"], res["achieved_tol"] = int(m.group(1)), float(m.group(2))
    m = re.search(r"IMLSolver info: user time consumed by solution: ([\d.]+)s", log)
    if m:
        res["solve_s"] = float(m.group(1))
    m = re.search(r"user time consumed by solution step 1: ([\d.]+)s", log)
    if m:
        res["solution_step_user_s"] = float(m.group(1))
    m = re.search(r"CudaTiming (\{.*\})", log)
    if m:
        res["phases"] = json.loads(m.group(1))
    if r.returncode:
        res["log_tail"] = log[-1500:]
    return res


def tip_deflection(outfile, node):
    txt = open(outfile).read()
    m = re.search(r"Node\s+%d .*?\n((?:\s+dof .*\n)+)" % node, txt)
    vals = re.findall(r"dof\s+3\s+d\s+([-+.\de]+)", m.group(1)) if m else []
    return float(vals[0]) if vals else None


def measure(dims=(125, 32, 32), lstol="1e-3", skip_reference=False):
    from bench import host_cores
    ours_exe = os.path.join(ROOT, "plugin", "_build", "oofem_cuda")
    ref_exe = next((p for p in (os.path.join(ROOT, "oracle", "_ref", n) for n in ("oofem_omp", "oofem")) if os.path.exists(p)), None)
    if not os.path.exists(ours_exe):
        return {"unavailable": "plugin/_build/oofem_cuda not built"}
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        fns, nelem, tipnode = write_case(td, dims, lstol)
        res = {"input": f"LinearStatic, {dims[0]}x{dims[1]}x{dims[2]} = {nelem} LSpace elements, IML CG + diagonal preconditioner to lstol {lstol}, "
                        f"output of one node and one element", "input_write_s": round(time.time() - t0, 2)}
        env = dict(os.environ, OOFEM_B200_TIMING="1")
        res["ours"] = run(ours_exe, fns["cuda"], env)
        res["ours"]["exe"] = "plugin/_build/oofem_cuda (reference main.C + cudacsr/cudacg plugin, lstype 9 smtype 11)"
        try:
            res["ours"]["tip_w"] = tip_deflection(os.path.join(td, "cuda.out"), tipnode)
        except Exception:
            pass
        if ref_exe and not skip_reference:
            env = dict(os.environ)
            env["OMP_NUM_THREADS"] = str(host_cores())
            res["reference"] = run(ref_exe, fns["ref"], env)
            res["reference"]["exe"] = os.path.relpath(ref_exe, ROOT) + f" (unmodified reference, lstype 1 smtype 2, {host_cores()} OpenMP threads)"
            try:
                res["reference"]["tip_w"] = tip_deflection(os.path.join(td, "ref.out"), tipnode)
            except Exception:
                pass
            if res["ours"]["rc"] == 0 and res["reference"]["rc"] == 0:
                res["wall_ratio"] = round(res["reference"]["wall_s"] / res["ours"]["wall_s"], 2)
    return res


if __name__ == "__main__":
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    dims = tuple(int(v) for v in a[:3]) if len(a) >= 3 else (125, 32, 32)
    lstol = sys.argv[sys.argv.index("--lstol") + 1] if "--lstol" in sys.argv else "1e-3"
    print(json.dumps(measure(dims, lstol, "--skip-reference" in sys.argv)))
