"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/msmgpu.h declares; without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

from newmsm_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def test_device_pow_self_test_on_this_host():
    """csrc/hostpow.cu: the pow tables of the C library mapped into the process are found and the host-side restatement of its algorithm
    agrees with std::pow on ~2 M arguments (bit for bit); a corrupted table entry is noticed and switches the device path off."""
    import subprocess, sys
    flags = open("/proc/cpuinfo").read()
    code = "from newmsm_b200 import capi; print('POW', capi.lib().msmgpu_device_pow_enabled())"
    run = lambda extra: subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True, cwd=ROOT).stdout
    if " fma " in flags and " avx2 " in flags:      # glibc selects the variant restated in hostpow.cuh on FMA + AVX2 hosts
        assert "POW 1" in run({})
    assert "POW 0" in run({"MSMGPU_POW_SELFTEST_CORRUPT": "1"})
    assert "POW 0" in run({"MSMGPU_DEVICE_POW": "0"})
