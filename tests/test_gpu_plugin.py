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
    """StaticStructural + NRSolver + MisesMat: the hook declines (not linear elastic), the reference's host
    loop assembles every Newton tangent into cudacsr and cudacg solves it."""
    need(DUMP)
    out = tmp_path / "dump.bin"
    r = subprocess.run([DUMP, cuda_input(name, tmp_path), str(out)], capture_output=True, text=True, cwd=tmp_path, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    d = read_dump(str(out))
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert d["meta"][3] == 0
    assert np.array_equal(d["rowptr"], g["colptr"]) and np.array_equal(d["colind"], g["rowind"])
    assert relerr(d["node_u"], g["node_u"]) < 1e-8
    # (the tangent AFTER the converged step is not compared with the reference dump: with tempKappa == kappa up
    # to round-off the loading/unloading branch of MisesMat::give3dMaterialStiffnessMatrix is decided by noise)
    assert relerr(d["val_hostloop"], d["val"]) < 1e-14            # two host-loop assemblies of the same state agree


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
    assert "CudaCG" in r.stdout + r.stderr
    ours = _numbers(tmp_path / (name + ".out"))
    os.rename(tmp_path / (name + ".out"), tmp_path / "cuda.out")
    r = subprocess.run([REF_EXE, "-f", cuda_input(name, tmp_path, keep=True)], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    ref = _numbers(tmp_path / (name + ".out"))
    assert ours.size == ref.size and ours.size > 500
    assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max()    # the .out prints ~5 significant digits
