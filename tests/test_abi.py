"""The C-ABI library loads and exports every symbol include/b2rt.h declares; without a CUDA
device every entry point fails loudly (there is no CPU fallback to fall into)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_product


def _declared():
    with open(os.path.join(ROOT, "include", "b2rt.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b2rt_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    prod = load_product()
    prod.build_all(verbose=False)
    L = ctypes.CDLL(prod.lib_path())
    names = _declared()
    assert len(names) >= 27
    for n in names:
        assert hasattr(L, n), "libb2rt.so does not export %s" % n
    assert sorted(prod.capi.SYMBOLS) == names       # the Python view binds exactly the header's surface


def test_no_oracle_in_product():
    """The product library must not link or name the oracle (parity claims depend on it)."""
    prod = load_product()
    with open(prod.lib_path(), "rb") as f:
        blob = f.read()
    assert b"liboracle" not in blob and b"libref_oracle" not in blob and b"oracle_trace" not in blob
    pkg = os.path.join(ROOT, "mini-opencl-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "liboracle" not in src and "oracle_lib" not in src and "rt_oracle" not in src, fn


def test_status_strings_follow_the_reference_table():
    L = load_product().lib()
    assert L.b2rt_status_string(0) == b"CL_SUCCESS"                  # CLutils.h:29-105
    assert L.b2rt_status_string(-1) == b"CL_DEVICE_NOT_FOUND"
    assert L.b2rt_status_string(-52) == b"CL_INVALID_KERNEL_ARGS"
    assert L.b2rt_status_string(-63) == b"CL_INVALID_GLOBAL_WORK_SIZE"
    assert L.b2rt_status_string(-999) == b"Unknown OpenCL error"


def test_fails_loudly_without_a_gpu():
    prod = load_product()
    if prod.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(prod.B2RTError) as e:
        prod.Context(0)
    assert e.value.status == -1 and "no CPU fallback" in str(e.value)
    L = prod.lib()
    assert L.b2rt_finish(None) == -34                                  # CL_INVALID_CONTEXT
    rays = np.zeros(4, dtype=prod.RAY_DTYPE)
    hits = np.zeros(4, dtype=prod.HIT_DTYPE)
    assert L.b2rt_trace_closest(None, rays.ctypes.data, 4, hits.ctypes.data) == -34
