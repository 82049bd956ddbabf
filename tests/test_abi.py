"""CPU-side checks of the drop-in boundary: libtfl.so loads and exports every symbol
include/tfl.h declares; without a CUDA device the product refuses to run (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "tfl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tfl_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from fluidnet_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()


def test_library_exports_every_symbol():
    from fluidnet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libtfl.so does not export " + s
    lib.tfl_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.tfl_version()
    assert lib.tfl_advect_method_from_string(b"maccormackOurs") == 5
    assert lib.tfl_advect_method_from_string(b"bogus") == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fluidnet_b200 import tfluids
    from fluidnet_b200._lib import TflError
    t = torch.zeros(1, 1, 4, 4, 4)
    with pytest.raises(TflError):
        tfluids.emptyDomain(t, True, 1)
    h = ctypes.c_void_p()
    from fluidnet_b200 import _lib
    assert _lib.load().tfl_create(ctypes.byref(h), 0) != 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under fluidnet_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fluidnet_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".lua")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "libtfluids_ref" not in txt, f
