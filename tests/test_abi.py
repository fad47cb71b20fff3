"""CPU: libcommpy_b200.so loads without a GPU and exports every symbol include/commpy_b200.h declares."""
import ctypes
import os
import re

from commpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "commpy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 20
    assert sorted(_lib.SYMBOLS) == names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_status_strings_no_gpu_needed():
    lib = _lib.load()
    assert lib.cpb_strerror(0) == b"ok"
    assert b"invalid" in lib.cpb_strerror(1)
    assert lib.cpb_version() >= 100


def test_no_cpu_fallback_in_product():
    """The product package must never import the oracle."""
    pkg = os.path.join(ROOT, "commpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "commpy_oracle" not in txt, f
