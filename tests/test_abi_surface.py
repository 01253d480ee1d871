"""CPU-only: the C-ABI library loads and exports every symbol include/cutesdr_cuda.h declares, the
Python prototypes cover the same set, and without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import cutesdr_b200 as cs
from cutesdr_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cutesdr_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cutesdr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_prototyped():
    syms = declared_symbols()
    assert len(syms) > 60
    lib = C.CDLL(cs.library_path())
    for s in syms:
        assert hasattr(lib, s), "libcutesdr_cuda.so does not export %s" % s
    assert sorted(L.PROTOTYPES) == syms, set(L.PROTOTYPES) ^ set(syms)
    cs.load_library()


def test_header_cites_reference_interface_for_every_entry_point():
    txt = open(os.path.join(ROOT, "include", "cutesdr_cuda.h")).read()
    assert txt.count("dsp/") > 40          # file:line citations of the replaced reference methods


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cs.CuteSdrError):
        cs.ReceiverBank(4, 2e6)
    with pytest.raises(cs.CuteSdrError):
        cs.CFft()
    with pytest.raises(cs.CuteSdrError):
        cs.CDownConvert()


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under cutesdr_b200/ may import, include or load it."""
    pkg = os.path.join(ROOT, "cutesdr_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s*[\"<][^\">]*oracle|liboracle|_ref/lib", re.M)
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), "%s reaches into oracle/" % os.path.join(dirpath, f)
