"""The C-ABI boundary without a GPU: include/intel_b200.h is valid C, libintel_b200.so loads, and it exports every
function the header declares (and the ctypes layer declares exactly those).  No compute entry point is called."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "intel_b200.h")


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(intel_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from intel_sigir2023_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from intel_sigir2023_b200.build import build
        build()
    return _lib.LIB_PATH


def test_header_is_plain_c():
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_library_exports_every_declared_function(lib_path):
    names = declared_functions()
    assert len(names) >= 40, names
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_layer_matches_the_header(lib_path):
    from intel_sigir2023_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared_functions()
    lib = _lib.load()                       # _declare() raises AttributeError when a signature names a missing symbol
    assert lib.intel_abi_version() == 1
    assert isinstance(lib.intel_last_error(), (bytes, type(None)))


def test_missing_library_fails_loudly(tmp_path):
    from intel_sigir2023_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.load(str(tmp_path / "libintel_b200.so"))
