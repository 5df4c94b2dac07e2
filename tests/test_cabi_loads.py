"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/msmgpu.h declares; without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

from newmsm_b200 import build, capi


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return capi.lib()


def test_exports_every_declared_symbol(lib):
    declared = capi.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(capi._SIGNATURES), set(declared) ^ set(capi._SIGNATURES)


def test_header_cites_reference_lines():
    text = open(capi.HEADER_PATH).read()
    assert len(re.findall(r"replaces:", text)) >= 15
    assert "octree.cpp:156-214" in text and "resampler.cpp:72-140" in text and "cpp:236-243" in text


def test_no_silent_cpu_fallback(lib):
    import ctypes as C
    if lib.msmgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = lib.msmgpu_ctx_create(0, None, C.byref(h))
    assert st == capi.ERR_CUDA and b"CUDA" in lib.msmgpu_last_error()


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "newmsm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synth.py", f"{f} mentions the oracle"
