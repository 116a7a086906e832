"""A C++ program written against the reference's host classes (tests/cpp/reference_main.cpp: global `eng`,
CLEngineBase::renderLoop, CLRaytracer::pixels) compiled against the host mirror (host/glaze3d.h, libglaze3d.so +
libb2rt.so). Without a GPU it must fail loudly with the reference's CLException text; on a B200 its frames must match
the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import scenes
from conftest import ROOT, load_product

PKG = os.path.join(ROOT, "mini-opencl-raytracer_b200")


def _build(tmp_dir):
    load_product().build_all(verbose=False)
    exe = os.path.join(tmp_dir, "reference_main")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "reference_main.cpp"), "-I", os.path.join(PKG, "host"),
                           "-I", os.path.join(ROOT, "include"), "-L", PKG, "-lglaze3d", "-lb2rt", "-Wl,-rpath," + PKG, "-o", exe])
    return exe


def test_cpp_program_builds_and_fails_loudly_without_a_gpu(tmp_scene_dir):
    exe = _build(tmp_scene_dir)
    if load_product().device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, scenes.CORNELL, "64", "48", "1", "2", os.path.join(tmp_scene_dir, "x.raw")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "CL_DEVICE_NOT_FOUND" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_program_renders_the_oracles_frames(tmp_scene_dir, cornell_ref):
    exe = _build(tmp_scene_dir)
    W, H, frames, bounces = 160, 120, 6, 4
    out = os.path.join(tmp_scene_dir, "cornell.raw")
    r = subprocess.run([exe, scenes.CORNELL, str(W), str(H), str(frames), str(bounces), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "frames %d" % frames in r.stdout
    img = np.fromfile(out, dtype=np.float32).reshape(-1, 4)
    tris, nodes, mats = cornell_ref
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in range(1, frames + 1):                       # CLRaytracer starts at m_FrameCount = 1 (CLRaytracer.h:30)
        ol.oracle_render(tris, nodes, mats, want, W, H, fc, bounces)
    assert scenes.psnr(img[:, :3], want[:, :3]) >= 50.0
    assert (img[:, :3] == want[:, :3]).all(axis=1).mean() > 0.98
