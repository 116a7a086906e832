"""The oracle (oracle/rt_oracle.c, our C restatement of kernel_bvh.cl) is pinned here against
(a) the committed golden vectors, which were produced by the reference's own sources compiled
verbatim (tests/golden/make_golden.py), and (b) that verbatim build itself (oracle/_ref) on
fresh seeded inputs. Bit-exact everywhere: same compiler flags, same libm."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import scenes

G = scenes.GOLDEN
MISS = 0xFFFFFFFF


def _ids(refhit):
    return np.where(refhit["hit"] != 0, refhit["tri"].astype(np.int64), -1)


def _hids(h):
    return np.where(h["tri"] == MISS, -1, h["tri"].astype(np.int64))


def test_golden_scene_matches_ref_build(cornell_ref):
    tris, nodes, mats = cornell_ref
    g = np.load(os.path.join(G, "cornell_scene.npz"))
    assert tris.shape == (72, 256) and nodes.shape == (39, 48) and mats.shape == (6, 64)
    assert np.array_equal(nodes.view(np.uint32).reshape(-1, 12)[:, 8], g["node_offset"])
    assert np.array_equal(nodes.view(np.float32).reshape(-1, 12)[:, [0, 1, 2, 4, 5, 6]], g["node_bounds"])
    assert np.array_equal(tris.view(np.uint32).reshape(-1, 64)[:, 60], g["tri_mtl"])


def test_oracle_hits_match_golden(cornell_ref):
    tris, nodes, _ = cornell_ref
    g = np.load(os.path.join(G, "cornell_hits.npz"))
    rays = np.ascontiguousarray(g["rays"]).view(ol.RAY).reshape(-1)
    h, attr = ol.oracle_closest(tris, nodes, rays, want_attr=True)
    want = np.where(g["hit"] != 0, g["tri"].astype(np.int64), -1)
    assert np.array_equal(_hids(h), want)
    assert np.array_equal(h["t"].view(np.uint32), g["t"].view(np.uint32))           # bit-exact t
    m = g["hit"] != 0
    assert np.array_equal(attr["pos"][m].view(np.uint32), g["pos"][m].view(np.uint32))
    assert np.array_equal(attr["normal"][m].view(np.uint32), g["normal"][m].view(np.uint32))
    assert m.mean() > 0.5


def test_oracle_frames_match_golden(cornell_ref):
    tris, nodes, mats = cornell_ref
    g = np.load(os.path.join(G, "cornell_frames.npz"))
    W, H = int(g["width"]), int(g["height"])
    img = np.zeros((W * H, 4), dtype=np.float32)
    ol.oracle_render(tris, nodes, mats, img, W, H, 1, 1)
    assert np.array_equal(img[:, :3].view(np.uint32), g["b1_f1"].view(np.uint32))
    for lt in (0, 1, 2):
        img = np.zeros((W * H, 4), dtype=np.float32)
        for fc in (1, 2, 3):
            ol.oracle_render(tris, nodes, mats, img, W, H, fc, 4, light_type=lt)
        assert np.array_equal(img[:, :3].view(np.uint32), g["b4_f123_lt%d" % lt].view(np.uint32)), lt
    img = np.zeros((W * H, 4), dtype=np.float32)
    ol.oracle_render(tris, nodes, mats, img, W, H, 0, 9)
    assert np.array_equal(img[:, :3].view(np.uint32), g["b9_f0"].view(np.uint32))


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built")
def test_oracle_matches_verbatim_reference_on_bumpy_mesh(bumpy_ref):
    tris, nodes, _ = bumpy_ref
    assert tris.shape[0] == 2 * 20480            # the loader emits every triangular face twice
    rays = scenes.shell_rays(60000, 10.0, seed=21)
    h = ol.oracle_closest(tris, nodes, rays)
    r = ol.ref_closest(tris, nodes, rays)
    assert np.array_equal(_hids(h), _ids(r))
    assert np.array_equal(h["t"].view(np.uint32), r["t"].view(np.uint32))
    # secondary rays: the reference accepts negative t
    b = scenes.bounce_rays(rays, h, scenes.tri_normals(tris, h), seed=22)
    hb = ol.oracle_closest(tris, nodes, b)
    rb = ol.ref_closest(tris, nodes, b)
    assert np.array_equal(_hids(hb), _ids(rb))
    assert np.array_equal(hb["t"].view(np.uint32), rb["t"].view(np.uint32))
    assert (hb["t"] < 0).mean() > 0.01           # the trap is actually exercised
    ax = scenes.axis_rays((-11, -11, -11), (11, 11, 11), 200, seed=23)
    assert np.array_equal(_hids(ol.oracle_closest(tris, nodes, ax)), _ids(ol.ref_closest(tris, nodes, ax)))


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built")
def test_oracle_frame_matches_verbatim_reference_on_bumpy_mesh(bumpy_ref):
    tris, nodes, mats = bumpy_ref
    W, H = 80, 60
    cam = dict(pos=(0.0, -30.0, 4.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
    a = np.zeros((W * H, 4), dtype=np.float32)
    b = np.zeros((W * H, 4), dtype=np.float32)
    for fc in (1, 2):
        ol.oracle_render(tris, nodes, mats, a, W, H, fc, 5, **cam)
        ol.ref_render(tris, nodes, mats, b, W, H, fc, 5, **cam)
    assert np.array_equal(a[:, :3].view(np.uint32), b[:, :3].view(np.uint32))


def test_any_hit_is_intersect_hit(bumpy_ref):
    """occluded = Intersect(tmax).hit; early exit must not change the boolean."""
    tris, nodes, _ = bumpy_ref
    rays = scenes.box_rays(20000, (-14, -14, -14), (14, 14, 14), seed=31)
    rays["tmax"][::3] = 6.0
    h = ol.oracle_closest(tris, nodes, rays)
    occ = ol.oracle_any(tris, nodes, rays)
    assert np.array_equal(occ != 0, h["tri"] != MISS)
    assert 0.05 < (occ != 0).mean() < 0.95


def test_threads_do_not_change_results(cornell_ref):
    tris, nodes, _ = cornell_ref
    rays = ol.oracle_camera_rays(64, 64, 3)
    a = ol.oracle_closest(tris, nodes, rays, threads=1)
    b = ol.oracle_closest(tris, nodes, rays, threads=5)
    assert a.tobytes() == b.tobytes()
