#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Needs oracle/_ref/oofem_dump (python oracle/build_ref.py; only possible where
/root/reference is mounted).  For each case an OOFEM input file is written with
oofem_b200.inputfile.write_input, the reference solves it, and oracle/ref_dump.cpp dumps
numbering, location arrays, the CompCol pattern of CompCol::buildInternalStructure, the
element matrices / internal forces of the reference's own LSpace / LTRSpace / material
code, assembled values, one CompCol::times product and the solution -- all float64.
The .in files are committed next to the .npz so that the fixtures can be regenerated
and inspected.

    python tests/golden/make_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oofem_b200 import meshgen                                    # noqa: E402
from oofem_b200.inputfile import DirichletBC, Material, NodalLoad, Problem, write_input  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DUMP = os.path.join(ROOT, "oracle", "_ref", "oofem_dump")


def read_dump(fn):
    b = open(fn, "rb").read()
    o, out = 0, {}
    while o < len(b):
        l, = struct.unpack_from("<i", b, o); o += 4
        name = b[o:o + l].decode(); o += l
        dt, n = struct.unpack_from("<iq", b, o); o += 12
        out[name] = np.frombuffer(b, dtype="<f8" if dt else "<i4", count=n, offset=o).copy()
        o += n * (8 if dt else 4)
    return out


def cantilever(etype, nx, ny, nz, jitter, mat, engng, params, tip_load=None, tip_disp=None, nsteps=1):
    gen = meshgen.hex_beam if etype == "lspace" else meshgen.tet_beam
    lx = float(nx) / ny
    coords, conn = gen(nx, ny, nz, lx, 1.0, 1.0)
    fixed, tip = meshgen.cantilever_bcs(coords, lx)
    if jitter:
        coords = meshgen.perturb(coords, jitter, seed=7)
    pb = Problem(title=f"{etype} cantilever {nx}x{ny}x{nz}", outfile="case.out", engng=engng,
                 params=dict(nsteps=nsteps, **params), coords=coords, elem_type=etype, conn=conn,
                 elem_mat=np.zeros(conn.shape[0], np.int32), materials=[mat])
    pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, fixed))
    pb.ltfs[1] = ("const", 1.0)
    if nsteps > 1:
        pb.ltfs[2] = ("pwl", np.array([0.0, float(nsteps)]), np.array([0.0, 1.0]))
    ramp = 2 if nsteps > 1 else 1
    if tip_load is not None:
        pb.loads.append(NodalLoad([1, 2, 3], list(tip_load), ramp, tip))
    if tip_disp is not None:
        pb.bcs.append(DirichletBC([3], [tip_disp], ramp, tip))
    return pb


IML = dict(lstype=1, smtype=2, lstol=1e-14, lsiter=20000, lsprecond=1)

CASES = {
    # BASELINE configs[0]: linear-elastic LSpace brick cantilever, LinearStatic + IML CG
    "lspace_cantilever": lambda: cantilever("lspace", 6, 2, 2, 0.06, Material("isole", 210.0e3, 0.3),
                                            "linearstatic", IML, tip_load=(0.0, 0.3, -1.0)),
    # same with a non-zero Dirichlet condition (internal-force path of linearstatic.C:226-231)
    "lspace_prescribed": lambda: cantilever("lspace", 4, 2, 3, 0.05, Material("isole", 30.0e3, 0.2),
                                            "linearstatic", IML, tip_disp=-0.01),
    # configs[2] in miniature: LTRSpace tetra mesh, linear elastic
    "ltrspace_cantilever": lambda: cantilever("ltrspace", 4, 2, 2, 0.05, Material("isole", 70.0e3, 0.33),
                                              "linearstatic", IML, tip_load=(0.1, 0.0, -0.5)),
    # configs[3] in miniature: MisesMat J2 plasticity, Newton-Raphson with the tangent
    # re-assembled every iteration (manrmsteps 1 -> nrsolverAccelNRM, src/core/nrsolver.C:122-125,294)
    "lspace_mises": lambda: cantilever("lspace", 6, 2, 2, 0.04,
                                       Material("misesmat", 210.0e3, 0.3, sig0=240.0, H=2100.0),
                                       "staticstructural", dict(rtolf=1e-10, maxiter=60, manrmsteps=1),
                                       tip_disp=-0.08, nsteps=4),
    "ltrspace_mises": lambda: cantilever("ltrspace", 4, 2, 2, 0.04,
                                         Material("misesmat", 210.0e3, 0.3, sig0=240.0, H=2100.0),
                                         "staticstructural", dict(rtolf=1e-10, maxiter=60, manrmsteps=1),
                                         tip_disp=-0.06, nsteps=3),
}


def main():
    if not os.path.exists(DUMP):
        sys.exit("oracle/_ref/oofem_dump missing: run `python oracle/build_ref.py` first")
    for name, mk in CASES.items():
        pb = mk()
        infile = os.path.join(HERE, name + ".in")
        write_input(infile, pb)
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "dump.bin")
            r = subprocess.run([DUMP, infile, out], cwd=td, capture_output=True, text=True)
            if r.returncode:
                print(r.stdout[-3000:], r.stderr[-3000:])
                sys.exit(f"reference failed on {name}")
            d = read_dump(out)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(f"{name}: neq={d['meta'][0]} nelem={d['meta'][2]} nnz={d['rowind'].size} "
              f"|u|max={np.abs(d['node_u']).max():.6e}")


if __name__ == "__main__":
    main()
