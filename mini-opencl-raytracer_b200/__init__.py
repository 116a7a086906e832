"""mini-opencl-raytracer_b200 -- B200 (sm_100a) replacement for the device side of
jstrom2002/Mini-OpenCL-Raytracer: BVH traversal + ray/triangle intersection of
kernel_bvh.cl behind the reference's host API.

The product is native: ``libb2rt.so`` (C ABI of ``include/b2rt.h`` + CUDA kernels)
and ``libglaze3d.so`` (C++ mirror of the reference's host classes). This Python
package is only a thin ctypes view of those two libraries for the test and bench
harnesses. It never imports anything under ``oracle/`` and has no CPU fallback:
every entry point raises when the CUDA library or a CUDA device is missing.

The directory name contains hyphens, so load it with ``importlib`` (see
``tests/conftest.py``) -- it registers itself as ``mor_b200``.
"""
from .capi import (B2RTError, Context, HIT_DTYPE, MISS, RAY_DTYPE, device_count, lib, lib_path)  # noqa: F401
from .build import build_all  # noqa: F401
from . import layouts  # noqa: F401
from . import host  # noqa: F401
from . import workloads  # noqa: F401


def __getattr__(name):
    # torch is only needed by the multi-GPU helpers: import them on first use
    if name == "sharding":
        import importlib
        return importlib.import_module(".sharding", __name__)
    raise AttributeError(name)
