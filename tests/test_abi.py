"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol
include/nlf_b200.h declares (no compute calls without a GPU)."""
import os
import re

import pytest

from neoloopfinder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nlf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nlf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.LIB_PATH):
        import neoloopfinder_b200.build as b
        b.build()
    L = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in nlf_b200.h but not exported"
    # the ctypes table binds exactly the declared interface
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert L.nlf_version() == 100


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle."""
    pkg = os.path.join(ROOT, "neoloopfinder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
