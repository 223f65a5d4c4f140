"""The C-ABI library loads on a CPU-only box and exports every symbol include/autobz_cuda.h declares
(no compute calls without a GPU); the product fails loudly without a device."""
import ctypes
import os
import re

import pytest

import autobz_b200 as ab
from autobz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "autobz_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abz_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.abz_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only boxes")
    with pytest.raises(_lib.AutoBZCudaError):
        _lib.Context(0)
    s = ab.FourierSeries(ab.synthetic.integer_lattice(3)[0][0, 0], period=1.0, lo=-1)
    prob = ab.IntegralProblem(ab.FourierIntegrand(ab.dos_integrand, s, 0.1, 0.0), ab.load_bz(ab.FBZ(3)))
    with pytest.raises(_lib.AutoBZCudaError):
        ab.solve(prob, ab.PTR(npt=4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "autobzcore.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import orc" not in txt and "liborc" not in txt and "oracle/" not in txt, f
