"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/oofem_b200.h declares; without a GPU the product fails loudly (no fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oofem_b200 import build as ob_build
from oofem_b200 import capi


@pytest.fixture(scope="module")
def libpath():
    return ob_build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "oofem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ob200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = set(re.findall(r"\bT (ob200_[a-z0-9_]+)", out))
    decl = declared_symbols()
    assert len(decl) >= 40
    missing = [s for s in decl if s not in exported]
    assert not missing, f"declared in include/oofem_b200.h but not exported: {missing}"


def test_ctypes_binding_covers_header(libpath):
    assert sorted(capi.SYMBOLS) == declared_symbols()
    L = capi.lib()                 # resolves every symbol (AttributeError otherwise)
    assert L.ob200_version().decode().startswith("oofem_b200")


def test_no_cpu_fallback_without_device(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.OofemB200Error) as ei:
        capi.Context(0)
    assert ei.value.code == capi.ENODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under oofem_b200/ may reference it."""
    pkg = os.path.join(ROOT, "oofem_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "liboofem_oracle" not in txt and "oofem_oracle.c" not in txt, fn
