"""CPU: the C-ABI library loads and exports every symbol include/b200bit.h declares (no compute calls)."""
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "b200bit.h")).read()
    return sorted(set(re.findall(r"B200BIT_API\s+[\w\s\*]+?\b(b200bit_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert "b200bit_mpq_forward" in names and "b200bit_last_error" in names


def test_library_exports_every_declared_symbol():
    from bitorch_engine_b200 import _cabi
    lib = _cabi.lib()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/b200bit.h but not exported"
    assert lib.b200bit_version() == 100


def test_python_prototypes_cover_the_header():
    from bitorch_engine_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == _declared()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under bitorch-engine_b200/ may import it."""
    pkg = os.path.join(ROOT, "bitorch-engine_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
