"""The C-ABI library builds, loads and exports every symbol include/nutsb200.h declares.
No compute here (this runs without a GPU); on a box without a device the library must
refuse to create a context rather than fall back to anything."""
import ctypes
import re
from pathlib import Path

import pytest

from nuts333_b200 import api, build

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    h = (ROOT / "include" / "nutsb200.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(nutsb_[a-z0-9_]+)\s*\(", h)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(build.build()))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.nutsb_version() == 0x000100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    lib = api.bind(ctypes.CDLL(str(build.build())))
    h = ctypes.c_void_p()
    assert lib.nutsb_create(ctypes.byref(h), 0) == api.E_CUDA and not h
    with pytest.raises(api.NutsbError):
        api.Context(0)


def test_generator_library_builds():
    assert build.build_gen().exists()
