"""Host-side logic that needs no GPU: mesh generation, numbering, the OOFEM input reader/writer."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oofem_b200 import meshgen
from oofem_b200.inputfile import (SMT_CUDACSR, ST_CUDACG, DirichletBC, Material, NodalLoad, Problem, read_input,
                                  write_input)


def test_hex_beam_positive_jacobians():
    coords, conn = meshgen.hex_beam(3, 2, 2)
    assert coords.shape == (4 * 3 * 3, 3) and conn.shape == (12, 8)
    c = coords[conn - 1]
    # top face (nodes 1-4) above bottom face (5-8), FEI3dHexaLin ordering
    assert np.all(c[:, :4, 2] > c[:, 4:, 2])
    assert np.allclose(c[:, 0, :2], c[:, 4, :2])


def test_tet_beam_volumes_fill_the_box():
    coords, conn = meshgen.tet_beam(2, 2, 2, 2.0, 1.0, 1.0)
    c = coords[conn - 1]
    d = c[:, 1:] - c[:, :1]
    vol = np.abs(np.linalg.det(d)) / 6.0
    assert np.isclose(vol.sum(), 2.0)
    assert conn.shape == (2 * 2 * 2 * 6, 4)


def test_equation_numbering_matches_reference_dump():
    pb, d = load_golden("lspace_prescribed")
    nodeeq, neq = meshgen.equation_numbers(pb.coords.shape[0], pb.fixed_mask())
    assert neq == d["meta"][0]
    assert np.array_equal(nodeeq.reshape(-1), d["node_eq"])
    assert np.array_equal(meshgen.location_arrays(pb.conn, nodeeq).reshape(-1), d["elem_loc"])


def test_input_roundtrip(tmp_path):
    coords, conn = meshgen.hex_beam(2, 1, 1)
    fixed, tip = meshgen.cantilever_bcs(coords, 2.0)
    pb = Problem(title="t", engng="linearstatic", params=dict(nsteps=1, lstype=ST_CUDACG, smtype=SMT_CUDACSR, lstol=1e-9),
                 coords=coords, elem_type="lspace", conn=conn, elem_mat=np.zeros(2, np.int32),
                 materials=[Material("isole", 10.0, 0.2)])
    pb.ltfs[1] = ("const", 1.0)
    pb.bcs.append(DirichletBC([1, 2, 3], [0, 0, 0], 1, fixed))
    pb.loads.append(NodalLoad([3], [-1.0], 1, tip))
    fn = tmp_path / "a.in"
    write_input(str(fn), pb)
    q = read_input(str(fn))
    assert q.engng == "linearstatic" and q.params["lstype"] == ST_CUDACG and q.params["smtype"] == SMT_CUDACSR
    assert np.array_equal(q.conn, conn) and np.allclose(q.coords, coords)
    assert np.array_equal(q.fixed_mask(), pb.fixed_mask())
    assert np.allclose(q.nodal_load_vector(1.0), pb.nodal_load_vector(1.0))


def test_solver_keywords_by_name(tmp_path):
    src = open(os.path.join(GOLDEN, "lspace_cantilever.in")).read()
    src = src.replace("lstype 1 smtype 2", "lstype cudacg smtype cudacsr")
    fn = tmp_path / "b.in"
    fn.write_text(src)
    q = read_input(str(fn))
    assert q.params["lstype"] == ST_CUDACG and q.params["smtype"] == SMT_CUDACSR


def test_unsupported_records_fail_loudly(tmp_path):
    src = open(os.path.join(GOLDEN, "lspace_cantilever.in")).read().replace("lspace ", "qspace ")
    fn = tmp_path / "c.in"
    fn.write_text(src)
    with pytest.raises(ValueError):
        read_input(str(fn))


def test_piecewise_linear_function():
    pb = Problem()
    pb.ltfs[2] = ("pwl", np.array([0.0, 4.0]), np.array([0.0, 1.0]))
    assert pb.ltf_value(2, 1.0) == 0.25 and pb.ltf_value(2, 9.0) == 1.0


def test_plugin_parallel_for_covers_every_index_once(tmp_path):
    """plugin/cudaparallel.h (the std::thread loops of the plugin's host side): every index of [0, n) is visited exactly once,
    for sizes around the grain and the thread count, with one thread and with several."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "pf.cpp"
    src.write_text(r'''
#include "cudaparallel.h"
#include <cstdio>
#include <vector>
int main() {
    const long sizes[] = { 0, 1, 7, 4095, 4096, 4097, 8192, 65537, 1000003 };
    for ( long n : sizes ) {
        std::vector< int > hit(n, 0), threads(64, 0);
        oofem::parallelFor(n, [ & ](long b, long e, int t) {
            if ( b > e || b < 0 || e > n || t < 0 || t >= oofem::cudaPluginThreads() ) std::abort();
            threads[t]++;
            for ( long i = b; i < e; i++ ) hit[i]++;
        });
        for ( long i = 0; i < n; i++ ) if ( hit[i] != 1 ) { std::printf("n=%ld index %ld hit %d times\n", n, i, hit[i]); return 1; }
        for ( int t : threads ) if ( t > 1 ) { std::printf("thread index used twice\n"); return 1; }
    }
    std::printf("ok %d\n", oofem::cudaPluginThreads());
    return 0;
}
''')
    exe = tmp_path / "pf"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "plugin"), str(src), "-o", str(exe), "-lpthread"])
    for nthreads in ("1", "3", "16"):
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, OOFEM_B200_THREADS=nthreads))
        assert r.returncode == 0 and r.stdout.strip() == "ok " + nthreads, (r.stdout, r.stderr)


@pytest.mark.parametrize("patch,target", [("engngm_hook.patch", "src/core/engngm.C"),
                                          ("structengngmodel_hook.patch", "src/sm/EngineeringModels/structengngmodel.C")])
def test_hook_patches_apply_to_the_reference(patch, target, tmp_path):
    """The hook sites INTEGRATION.md shows are what plugin/build_plugin.py patches into scratch copies of the reference's files;
    each patch must apply cleanly and only add lines (plus the loop-bound rename of the vector hook)."""
    import shutil
    import subprocess
    ref = os.environ.get("OOFEM_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box): the prebuilt plugin/_build travels instead")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    work = tmp_path / os.path.basename(target)
    shutil.copy(os.path.join(ref, target), work)
    r = subprocess.run(["patch", "-s", str(work), os.path.join(root, "plugin", patch)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    before = open(os.path.join(ref, target)).read().splitlines()
    after = open(work).read().splitlines()
    assert 0 < len(after) - len(before) <= 12
    assert any("batched" in l.lower() for l in after) and not any("batched" in l.lower() for l in before)
