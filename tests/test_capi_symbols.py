"""The C-ABI library loads and exports every symbol include/pmm_b200.h declares.
No compute call is made (no GPU here); pmm_create must fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pmm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmm_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from laml_b200 import capi
    assert sorted(capi.SYMBOLS) == declared_symbols()


def test_library_exports_all_symbols():
    import __graft_entry__ as g
    g.build()
    from laml_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.pmm_abi_version() == 2


def test_no_cpu_fallback():
    """without a CUDA device the engine must refuse, not silently compute elsewhere"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from laml_b200.capi import Engine, PMMError
    with pytest.raises(PMMError, match="no usable CUDA device|CUDA"):
        Engine(2, np.array([0, 1], np.int32), np.array([1.0]), np.zeros((2, 1), np.int32))


def test_product_does_not_import_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "laml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s references the oracle" % f
