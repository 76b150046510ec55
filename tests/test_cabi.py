"""CPU checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports
every symbol include/mclst_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import subprocess

import pytest

from mclstexp_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_every_header_symbol_is_exported(lib):
    names = _lib.header_symbols()
    assert "mclst_find_matches" in names and "mclst_weighted_average" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.mclst_version() >= 100


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, out


def test_argument_validation_needs_no_gpu(lib):
    n = ctypes.c_size_t()
    assert lib.mclst_find_matches_workspace_bytes(1000, 10, 256, 50, 0, ctypes.byref(n)) == 0
    assert n.value > 0
    assert lib.mclst_find_matches_workspace_bytes(1000, 10, 0, 50, 0, ctypes.byref(n)) == -1
    assert b"bad shape" in lib.mclst_last_error()
    # null pointers are rejected before anything touches the device
    rc = lib.mclst_find_matches(None, 10, 256, None, 1, 256, 256, 5, 0, None, None, None, 0, 0, None)
    assert rc == -1


def test_product_has_no_cpu_fallback():
    import numpy as np
    import torch
    from mclstexp_b200 import retrieval
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MclstError):
        retrieval.find_matches(np.zeros((4, 8), np.float32), np.zeros((2, 8), np.float32), 1)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "mclstexp_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
