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
