"""pytest configuration: the `gpu` marker, package loading and session-wide scene fixtures."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def load_product():
    """Import mini-opencl-raytracer_b200/ (hyphenated directory) as module `mor_b200`."""
    if "mor_b200" in sys.modules:
        return sys.modules["mor_b200"]
    pkg = os.path.join(ROOT, "mini-opencl-raytracer_b200")
    spec = importlib.util.spec_from_file_location("mor_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mor_b200"] = mod
    spec.loader.exec_module(mod)
    if not (os.path.exists(mod.lib_path()) and os.path.exists(mod.host.lib_path())):
        mod.build_all(verbose=False)          # a checkout without built artefacts: compile in-tree (nvcc needs no GPU)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def product():
    return load_product()


@pytest.fixture(scope="session")
def tmp_scene_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenes"))


def _reference_scene(path, max_prims=4):
    """Scene arrays from the reference's own loader + builder (oracle/_ref). Where that build is not available (a
    checkout without /root/reference and without the prebuilt .so) the product's host mirror stands in: it produces
    the identical arrays (tests/test_host_scene.py proves that wherever oracle/_ref exists)."""
    import oracle_lib
    if oracle_lib.ref() is not None:
        return oracle_lib.ref_load_scene(path, max_prims)
    return load_product().host.load_scene(path, max_prims)


@pytest.fixture(scope="session")
def cornell_ref():
    """cornell.obj ingested by the reference's own loader + BVH builder (oracle/_ref)."""
    import oracle_lib
    import scenes
    return _reference_scene(scenes.CORNELL)


@pytest.fixture(scope="session")
def bumpy_ref(tmp_scene_dir):
    """Displaced icosphere, 20480 faces -> 40960 CLTriangle, via the reference's loader + builder."""
    import oracle_lib
    import scenes
    p, n, f = scenes.displaced_sphere(5)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "bumpy5.obj"), p, n, f)
    return _reference_scene(path)


@pytest.fixture(scope="session")
def gpu_ctx(product):
    ctx = product.Context(0)
    yield ctx
    ctx.close()
