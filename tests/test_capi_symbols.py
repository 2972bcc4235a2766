"""The C-ABI library loads and exports every symbol include/cfmm_b200.h
declares; without a GPU every compute entry point fails loudly (no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "cfmm_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfmm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cr):
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(cr.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cfmm_b200.h but not exported"
    # and the Python binding covers exactly the declared set
    from cfmmrouter_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_no_cpu_fallback(cr):
    lib = cr.load_library()
    assert lib.cfmm_version().decode().count(".") == 2
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure check is for GPU-less hosts")
    with pytest.raises(cr.CFMMError) as e:
        cr.DevicePools(4)
    assert e.value.code == -2 and "no CPU path" in e.value.message


def test_no_oracle_in_product_path():
    """The product never imports, links or calls the oracle."""
    pkg = os.path.join(ROOT, "cfmmrouter.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "liboracle" not in text and "oracle_" not in text and "import oracle" not in text, fn
    out = os.popen(f"ldd {os.path.join(pkg, 'libcfmm_b200.so')}").read()
    assert "oracle" not in out
