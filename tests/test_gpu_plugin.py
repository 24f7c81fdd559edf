"""The C++ plugin (plugin/*.C) inside the reference: OOFEM's own executable and engineering models,
built with the two new types, run on the committed golden inputs with `lstype 9 smtype 11`.

plugin/_build/oofem_dump_cuda (plugin/plugin_dump.cpp over the reference's objects + the plugin) writes
full-precision dumps; they are compared with the dumps of the UNMODIFIED reference (tests/golden/*.npz):
CSR integers bit-exact, assembled values / SpMV 1e-12, displacements 1e-8 (north_star).  The stock
executable plugin/_build/oofem_cuda is run on the same input and its .out compared with the .out of
oracle/_ref/oofem (the reference binary, which travels to the GPU box)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conftest import GOLDEN, relerr  # noqa: E402


def read_dump(fn):
    """Records of plugin/plugin_dump.cpp: [int32 namelen][name][int32 dtype][int64 count][payload]."""
    import struct
    b = open(fn, "rb").read()
    o, out = 0, {}
    while o < len(b):
        l, = struct.unpack_from("<i", b, o)
        o += 4
        name = b[o:o + l].decode()
        o += l
        dt, n = struct.unpack_from("<iq", b, o)
        o += 12
        out[name] = np.frombuffer(b, dtype="<f8" if dt else "<i4", count=n, offset=o).copy()
        o += n * (8 if dt else 4)
    return out

BUILD = os.path.join(ROOT, "plugin", "_build")
DUMP = os.path.join(BUILD, "oofem_dump_cuda")
EXE = os.path.join(BUILD, "oofem_cuda")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "oofem")

pytestmark = pytest.mark.gpu


def cuda_input(name, tmp_path, keep=False):
    """The committed golden input with the solver / matrix keywords switched to the new types."""
    lines = open(os.path.join(GOLDEN, name + ".in")).read().splitlines()
    lines[0] = str(tmp_path / (name + ".out"))
    rec = lines[2]
    if not keep:
        if "lstype" in rec:
            rec = rec.replace("lstype 1 smtype 2", "lstype 9 smtype 11")
        else:                                                    # StaticStructural inputs use the defaults
            rec = rec.replace(" nmodules", " lstype 9 smtype 11 lstol 1e-14 lsiter 20000 lsprecond 1 nmodules")
        assert "lstype 9 smtype 11" in rec
    lines[2] = rec
    fn = tmp_path / (name + ("_ref.in" if keep else "_cuda.in"))
    fn.write_text("\n".join(lines) + "\n")
    return str(fn)


def need(path):
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run python plugin/build_plugin.py (needs the reference tree) before shipping to the GPU box")


@pytest.mark.parametrize("name", ["lspace_cantilever", "ltrspace_cantilever", "lspace_prescribed"])
def test_plugin_linear_static_vs_reference_dump(name, tmp_path):
    need(DUMP)
    out = tmp_path / "dump.bin"
    r = subprocess.run([DUMP, cuda_input(name, tmp_path), str(out)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    d = read_dump(str(out))
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert d["meta"][3] == 1, "the batched element-evaluation hook was not taken"
    # pattern symmetric: CSR rowptr/colind are CompCol's colptr/rowind, bit for bit
    assert np.array_equal(d["rowptr"], g["colptr"]) and np.array_equal(d["colind"], g["rowind"])
    assert relerr(d["val"], g["val"]) < 1e-12                     # batched hook (GPU element kernels)
    assert relerr(d["val_hostloop"], g["val"]) < 1e-12            # host loop -> CudaCSR::assemble(loc, mat)
    assert np.array_equal(d["spmv_x"], g["spmv_x"]) and relerr(d["spmv_y"], g["spmv_y"]) < 1e-12
    assert relerr(d["node_u"], g["node_u"]) < 1e-8                # LinearStatic through cudacg
    a11, a11p, a12p, v1 = d["at_probe"]                           # SparseMtrx::at read / write-through
    assert a11 == d["val_hostloop"][0] and a11p == a11 + 1.0 and a12p == v1 - 0.5


@pytest.mark.parametrize("name", ["lspace_mises", "ltrspace_mises"])
def test_plugin_newton_raphson_mises_vs_reference_dump(name, tmp_path):
    """StaticStructural + NRSolver + MisesMat entirely through the hooks: every Newton tangent
    (EngngModel::assemble), every internal-force vector with its element norms
    (EngngModel::assembleVectorFromElements) and the status commit (EngngModel::updateYourself) run on the GPU;
    the displacements match the unmodified reference to 1e-8."""
    need(DUMP)
    out = tmp_path / "dump.bin"
    r = subprocess.run([DUMP, cuda_input(name, tmp_path), str(out)], capture_output=True, text=True, cwd=tmp_path, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    d = read_dump(str(out))
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert d["meta"][3] == 1, "the batched matrix hook was not taken for MisesMat"
    assert d["meta"][4] > 0, "the batched internal-force hook was not taken"
    assert d["meta"][5] > 0, "the batched status update did not run"
    assert np.array_equal(d["rowptr"], g["colptr"]) and np.array_equal(d["colind"], g["rowind"])
    assert relerr(d["node_u"], g["node_u"]) < 1e-8
    # (the tangent AFTER the converged step is not compared: with tempKappa == kappa up to round-off the
    # loading/unloading branch of MisesMat::give3dMaterialStiffnessMatrix is decided by noise)

    # the same input with the hooks switched off: host loops into cudacsr / cudacg
    out2 = tmp_path / "dump_host.bin"
    r = subprocess.run([DUMP, cuda_input(name, tmp_path), str(out2)], capture_output=True, text=True, cwd=tmp_path, timeout=900,
                       env=dict(os.environ, OOFEM_B200_NO_BATCH="1"))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    h = read_dump(str(out2))
    assert h["meta"][3] == 0 and h["meta"][4] == 0
    assert relerr(h["node_u"], g["node_u"]) < 1e-8


def _numbers(fn):
    txt = open(fn).read()
    txt = txt[txt.index("Output for time"):]                       # skip the header (file names, timestamps)
    txt = re.sub(r"(User time consumed|Real time consumed|Total .* time).*", "", txt)
    return np.array([float(t) for t in re.findall(r"[-+]?\d+\.\d+e[-+]\d+", txt)])


def test_stock_executable_output_matches_reference_executable(tmp_path):
    need(EXE)
    need(REF_EXE)
    name = "lspace_cantilever"
    r = subprocess.run([EXE, "-f", cuda_input(name, tmp_path)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    log = r.stdout + r.stderr
    assert "CudaCG" in log
    # every per-element host loop of the step is taken by a hook: tangent, internal forces, the strain / stress update of
    # StructuralEngngModel::updateInternalState and the internal forces of computeReaction (prescribed numbering)
    for what in ("tangent assembly", "internal forces", "internal state update", "reaction forces"):
        assert f"batched {what} on the GPU" in log, (what, log[-3000:])
    ours = _numbers(tmp_path / (name + ".out"))
    os.rename(tmp_path / (name + ".out"), tmp_path / "cuda.out")
    r = subprocess.run([REF_EXE, "-f", cuda_input(name, tmp_path, keep=True)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    ref = _numbers(tmp_path / (name + ".out"))
    assert ours.size == ref.size and ours.size > 500
    assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max()    # the .out prints ~5 significant digits


def test_temperature_load_keeps_the_host_loops(tmp_path):
    """lspace_cantilever with a StructTemperatureLoad on every element (tests/golden/lspace_thermal.in): the thermal strain
    enters the host's stresses and internal forces through computeStressIndependentStrainVector_3d
    (structuralmaterial.C:2268-2340), which the resident kernels do not evaluate -- every batched hook must decline, the
    element loops run on the host into cudacsr / cudacg, and the .out equals the reference executable's."""
    need(EXE)
    need(REF_EXE)
    name = "lspace_thermal"
    r = subprocess.run([EXE, "-f", cuda_input(name, tmp_path)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-3000:]
    assert "CudaCG" in log and "batched" not in log, log[-3000:]
    ours = _numbers(tmp_path / (name + ".out"))
    os.rename(tmp_path / (name + ".out"), tmp_path / "cuda.out")
    r = subprocess.run([REF_EXE, "-f", cuda_input(name, tmp_path, keep=True)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    ref = _numbers(tmp_path / (name + ".out"))
    assert ours.size == ref.size and ours.size > 500
    assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max()


def _element_output(fn):
    """Numbers of the element records (strains, stresses, status variables per Gauss point) of an .out file."""
    txt = open(fn).read()
    txt = txt[txt.index("Element output:"):]
    txt = txt[:txt.index("R E A C T I O N S") if "R E A C T I O N S" in txt else len(txt)]
    return np.array([float(t) for t in re.findall(r"[-+]?\d+\.\d+e[-+]\d+", txt)])


@pytest.mark.parametrize("name", ["lspace_mises"])
def test_stock_executable_mises_output_matches_reference_executable(name, tmp_path):
    """oofem_cuda -f on the MisesMat Newton-Raphson input: the .out file (displacements, and the Gauss-point strains,
    stresses and plastic variables that the status update copies back from HBM) against the reference executable."""
    need(EXE)
    need(REF_EXE)
    r = subprocess.run([EXE, "-f", cuda_input(name, tmp_path)], capture_output=True, text=True, cwd=tmp_path, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    os.rename(tmp_path / (name + ".out"), tmp_path / "cuda.out")
    r = subprocess.run([REF_EXE, "-f", cuda_input(name, tmp_path, keep=True)], capture_output=True, text=True, cwd=tmp_path, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    ours, ref = _numbers(tmp_path / "cuda.out"), _numbers(tmp_path / (name + ".out"))
    assert ours.size == ref.size and ours.size > 500
    assert np.abs(ours - ref).max() <= 2e-5 * np.abs(ref).max()
    eo, er = _element_output(tmp_path / "cuda.out"), _element_output(tmp_path / (name + ".out"))
    assert eo.size == er.size and eo.size > 200 and np.abs(er).max() > 0
    assert np.abs(eo - er).max() <= 2e-5 * np.abs(er).max()


@pytest.mark.parametrize("name,hooked", [("deadweight02", True), ("patch300", True), ("patch301", True), ("Mises01", False),
                                         ("interface3d", None), ("InterfaceEL_SurfQuad1", None)])
def test_reference_own_tests_sm_through_the_plugin(name, hooked, tmp_path):
    """The reference's own tests/sm inputs (copied unchanged to tests/golden/ref_sm) with `lstype 9 smtype 11`: their
    #%BEGIN_CHECK% blocks are checked by the reference's errorcheck module inside the run (a mismatch is an OOFEM_ERROR and
    a non-zero exit).  deadweight02 (LSpace, dead weight, nodes with single prescribed dofs), patch300 and patch301 (LTRSpace)
    take the hooks; Mises01 is a truss1d model with both ends prescribed (no equation): host loops only.  hooked = None: models the
    batched hooks must DECLINE -- LTRSpace / LSpace mixed with interface elements (interface3d, a LinearStatic;
    InterfaceEL_SurfQuad1) -- so that every element matrix reaches the cudacsr matrix through SparseMtrx::assemble(loc, mat) and
    cudacg solves.  (brick_nlgeo_1 and tutorialmaterial, also LSpace-only inputs of tests/sm, have no free dof: nothing to solve.)"""
    need(EXE)
    lines = open(os.path.join(GOLDEN, "ref_sm", name + ".in")).read().splitlines()
    k = next(i for i, l in enumerate(lines) if l.lower().startswith(("staticstructural", "linearstatic")))
    lines[k] = lines[k].replace(" nmodules", " lstype 9 smtype 11 lstol 1e-14 lsiter 20000 lsprecond 1 nmodules", 1)
    fn = tmp_path / (name + ".in")
    fn.write_text("\n".join(lines) + "\n")
    r = subprocess.run([EXE, "-f", str(fn)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    log = r.stdout + r.stderr
    assert r.returncode == 0 and "Checking rules" in log and "Total 0 error(s)" in log, log[-3000:]
    if hooked is None:
        assert "CudaCG" in log and "batched tangent assembly on the GPU" not in log, log[-3000:]
        return
    assert ("CudaCG" in log) == hooked          # Mises01 has no free dof: nothing to solve, the check values still hold
    assert ("batched tangent assembly on the GPU" in log) == hooked, log[-3000:]
    assert ("batched internal forces on the GPU" in log) == hooked, log[-3000:]
    # the #REACTION values of the CHECK blocks come from computeReaction: internal forces in the prescribed numbering
    assert ("batched reaction forces on the GPU" in log) == hooked, log[-3000:]
