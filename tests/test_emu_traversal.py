"""CPU checks of the PRODUCT's traversal logic: csrc/traverse.cuh and csrc/wide_bvh.cpp compile
as plain C++ (tests/emu/, test only), so the compressed wide BVH + reference-order walk can be
compared with the oracle without a GPU. The CUDA build of the same text is what the -m gpu tests
exercise through the C ABI."""
import numpy as np

import oracle_lib as ol
import scenes
from conftest import _reference_scene

MISS = 0xFFFFFFFF


def _same(e, h):
    assert np.array_equal(e["tri"], h["tri"])
    assert np.array_equal(e["t"].view(np.uint32), h["t"].view(np.uint32))
    m = h["tri"] != MISS
    assert np.array_equal(e["u"][m].view(np.uint32), h["u"][m].view(np.uint32))
    assert np.array_equal(e["v"][m].view(np.uint32), h["v"][m].view(np.uint32))


def test_cornell_camera_and_shell(cornell_ref):
    tris, nodes, _ = cornell_ref
    st = ol.emu_build(tris, nodes)
    assert st.n_leaf_blocks == 20 and st.max_depth_binary >= 5
    rays = np.concatenate([ol.oracle_camera_rays(160, 120, 1),
                           scenes.shell_rays(20000, 12.0, seed=5, centre=(0.0, 7.0, 8.0)),
                           scenes.box_rays(20000, (-9, -2, -1), (9, 16, 17), seed=6),
                           scenes.axis_rays((-9, -2, -1), (9, 16, 17), 100, seed=7)])
    st2 = ol.EmuStats()
    _same(ol.emu_trace(rays, stats=st2), ol.oracle_closest(tris, nodes, rays))
    assert st2.overflow == 0
    assert np.array_equal(ol.emu_trace(rays, any_hit=True) != 0, ol.oracle_any(tris, nodes, rays) != 0)


def test_bumpy_primary_bounce_and_any(bumpy_ref):
    tris, nodes, _ = bumpy_ref
    st = ol.emu_build(tris, nodes)
    assert st.n_wide < nodes.shape[0] / 3
    rays = scenes.shell_rays(40000, 10.0, seed=41)
    h = ol.oracle_closest(tris, nodes, rays)
    _same(ol.emu_trace(rays), h)
    b = scenes.bounce_rays(rays, h, scenes.tri_normals(tris, h), seed=42)
    hb = ol.oracle_closest(tris, nodes, b)
    assert (hb["t"] < 0).mean() > 0.01
    _same(ol.emu_trace(b), hb)                               # negative-t trap: order-faithful walk
    short = b.copy()
    short["tmax"] = 3.0
    _same(ol.emu_trace(short), ol.oracle_closest(tris, nodes, short))
    assert np.array_equal(ol.emu_trace(short, any_hit=True) != 0, ol.oracle_any(tris, nodes, short) != 0)
    inside = scenes.box_rays(20000, (-5, -5, -5), (5, 5, 5), seed=43)   # culling: mostly misses from inside
    _same(ol.emu_trace(inside), ol.oracle_closest(tris, nodes, inside))


def test_stack_in_shared_columns_and_branch_free_push(bumpy_ref):
    """The kernels keep the first entries of a lane's stack in a strided shared-memory column (HybridStack) and, outside the
    counting build, push without a branch: same hits for every split between 'shared' and local entries, under any
    interleaving of node and leaf steps."""
    tris, nodes, _ = bumpy_ref
    ol.emu_build(tris, nodes)
    rays = scenes.shell_rays(20000, 10.0, seed=71)
    h = ol.oracle_closest(tris, nodes, rays)
    b = scenes.bounce_rays(rays, h, scenes.tri_normals(tris, h), seed=72)
    hb = ol.oracle_closest(tris, nodes, b)
    short = b.copy()
    short["tmax"] = 3.0
    occ = ol.oracle_any(tris, nodes, short) != 0
    for depth in (1, 2, 3, 12):
        for schedule in (0, 977 + depth):
            _same(ol.emu_trace_hybrid(rays, depth=depth, schedule=schedule), h)
            _same(ol.emu_trace_hybrid(b, depth=depth, schedule=schedule), hb)
            assert np.array_equal(ol.emu_trace_hybrid(short, any_hit=True, depth=depth, schedule=schedule) != 0, occ)


def test_wide_nodes_fetch_fewer_bytes_than_reference_layout(bumpy_ref):
    tris, nodes, _ = bumpy_ref
    ol.emu_build(tris, nodes)
    rays = scenes.shell_rays(5000, 10.0, seed=44)
    st = ol.EmuStats()
    ol.emu_trace(rays, stats=st)
    _, cnt = ol.oracle_closest(tris, nodes, rays, want_counters=True)
    ref_bytes = cnt["nodes_visited"] * 48 + cnt["tris_tested"] * 256
    assert st.words * 16 < 0.5 * ref_bytes
    assert st.tri_tests == cnt["tris_tested"]               # identical set of exact triangle tests


def test_degenerate_scenes(tmp_scene_dir):
    import os
    # one triangle (root is a leaf) and a quad face (the loader fans it into 3 overlapping triangles)
    p = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float32)
    n = np.tile(np.array([[0, 0, 1]], dtype=np.float32), (4, 1))
    for name, faces, quads in (("one.obj", [[0, 1, 2]], None), ("quad.obj", [[0, 1, 2]], [[0, 1, 3, 2]])):
        path = scenes.write_obj(os.path.join(tmp_scene_dir, name), p, n, np.array(faces), quads)
        tris, nodes, _ = _reference_scene(path)
        ol.emu_build(tris, nodes)
        rays = scenes.box_rays(4000, (-1, -1, 0.5), (2, 2, 3), seed=45)
        _same(ol.emu_trace(rays), ol.oracle_closest(tris, nodes, rays))


def test_exact_ties_across_leaves_follow_the_reference_order(tmp_scene_dir):
    """Coincident lattices: most hits are bit-exact t ties between triangles of different leaves; the winner is decided
    by the reference's visiting order alone. Every interleaving of node and leaf steps must reproduce it."""
    import os
    p, n, f = scenes.tie_grid(24, layers=2)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "ties.obj"), p, n, f)
    tris, nodes, _ = _reference_scene(path)
    ol.emu_build(tris, nodes)
    rays = scenes.tie_rays(24)
    want, cnt = ol.oracle_closest(tris, nodes, rays, want_counters=True)
    assert (want["tri"] != MISS).mean() > 0.9
    # the scene really is a tie scene: the oracle's winner is NOT simply the lowest index among equal-t candidates
    if ol.ref() is not None:                                  # the verbatim reference Intersect() agrees with the restatement here
        ref = ol.ref_closest(tris, nodes, rays)
        hit = ref["hit"] != 0
        assert np.array_equal(hit, want["tri"] != MISS)
        assert np.array_equal(ref["tri"][hit].astype(np.uint32), want["tri"][hit])
        assert np.array_equal(ref["t"][hit].view(np.uint32), want["t"][hit].view(np.uint32))
    # ... and it really is a tie scene: for many rays another triangle of ANOTHER leaf has the bit-identical t
    leaf_of = np.zeros(tris.shape[0], dtype=np.int64)
    nv = nodes.view(np.uint32).reshape(-1, 12)
    npr = nodes.view(np.uint16).reshape(-1, 24)[:, 18]
    for i in np.nonzero(npr > 0)[0]:
        leaf_of[nv[i, 8]:nv[i, 8] + npr[i]] = i
    sub = np.nonzero(want["tri"] != MISS)[0][::13][:300]
    ok, t_all = scenes.moller_trumbore_all(tris, rays[sub])
    win = want["tri"][sub].astype(np.int64)
    assert ok[np.arange(sub.size), win].all()
    assert np.array_equal(t_all[np.arange(sub.size), win].view(np.uint32), want["t"][sub].view(np.uint32))
    tied = ok & (t_all.view(np.uint32) == want["t"][sub].view(np.uint32)[:, None])         # every candidate with the winner's exact t
    leaves_tied = np.array([np.unique(leaf_of[np.nonzero(row)[0]]).size for row in tied])
    assert (leaves_tied >= 2).mean() > 0.5                     # most winners beat an equal-t candidate of ANOTHER leaf ...
    lowest = np.array([np.nonzero(row)[0].min() for row in tied])
    assert (lowest != win).mean() > 0.05                        # ... and not by having the lowest index: by being visited first
    for schedule in (0, 1, 12345):
        _same(ol.emu_trace(rays, schedule=schedule), want)
    assert np.array_equal(ol.emu_trace(rays, any_hit=True) != 0, ol.oracle_any(tris, nodes, rays) != 0)


def test_random_scenes_fuzz(tmp_scene_dir):
    """Small random scenes of awkward shapes (slivers, clusters, huge + tiny triangles, axis-aligned sheets, duplicated
    vertices, zero-area faces) with several maxPrimitivesInNode values: the product's wide-BVH builder + traversal state
    machine, under random step interleavings, must reproduce the oracle's walk over the reference-format arrays bit for
    bit -- closest hits, any-hit, and rays that start inside / on the geometry."""
    import os
    prod_host = __import__("conftest").load_product().host
    rng = np.random.default_rng(2024)
    for case in range(12):
        pos, faces = scenes.fuzz_scene(case, rng)
        nrm = np.tile(np.array([[0.0, 0.0, 1.0]], dtype=np.float32), (pos.shape[0], 1))
        path = scenes.write_obj(os.path.join(tmp_scene_dir, "fuzz%d.obj" % case), pos, nrm, faces)
        max_prims = int(rng.choice([1, 2, 4, 8, 64]))
        tris, nodes, _ = prod_host.load_scene(path, max_prims)
        if ol.ref() is not None:                                   # same arrays as the reference's own loader + builder
            rt, rn, _ = ol.ref_load_scene(path, max_prims)
            assert np.array_equal(rn.view(np.uint32).reshape(-1, 12)[:, 8], nodes.view(np.uint32).reshape(-1, 12)[:, 8])
        bound = ol.emu_build(tris, nodes).stack_bound
        lo, hi = pos.min(0) - 1.0, pos.max(0) + 1.0
        rays = np.concatenate([scenes.box_rays(1500, lo, hi, seed=case), scenes.axis_rays(lo, hi, 20, seed=case),
                               scenes.pack_rays(pos[rng.integers(0, pos.shape[0], 300)], rng.normal(size=(300, 3)))])   # origins ON vertices
        rays["tmax"][::5] = rng.uniform(0.1, 20.0, size=rays["tmax"][::5].shape).astype(np.float32)
        want = ol.oracle_closest(tris, nodes, rays)
        for schedule in (0, 7 + case):
            st = ol.EmuStats()
            _same(ol.emu_trace(rays, stats=st, schedule=schedule), want)
            assert st.overflow == 0 and st.max_stack <= bound          # the exact stack bound of the wide tree holds
        assert np.array_equal(ol.emu_trace(rays, any_hit=True, schedule=99) != 0, ol.oracle_any(tris, nodes, rays) != 0)


# ---- warp-cooperative tail mode (csrc/coop.cuh) on 32 emulated lanes (tests/emu/warp_emu.cpp) ----------------------
def _coop_same(rays, want, occluded=None):
    """Whole rays cooperatively from the root, hand-overs after a few / many solo steps, one node per round (the
    depth-first fallback) and four: all must give the oracle's hit, bit for bit."""
    for handoff, wide_limit, resume in ((0, 192, 0), (6, 192, 0), (60, 192, 0), (0, 0, 0), (9, 0, 0), (5, 192, 7), (30, 192, 40), (1, 0, 3)):
        st = ol.EmuStats()
        _same(ol.emu_trace_coop(rays, stats=st, handoff=handoff, wide_limit=wide_limit, resume=resume), want)   # resume: two-step tail (Lane::resume)
        assert st.overflow == 0
    if occluded is not None:
        for handoff, resume in ((0, 0), (7, 0), (7, 9)):
            assert np.array_equal(ol.emu_trace_coop(rays, any_hit=True, handoff=handoff, resume=resume) != 0, occluded != 0)


def test_coop_cornell(cornell_ref):
    tris, nodes, _ = cornell_ref
    ol.emu_build(tris, nodes)
    rays = np.concatenate([ol.oracle_camera_rays(80, 60, 1), scenes.shell_rays(2000, 12.0, seed=5, centre=(0.0, 7.0, 8.0)),
                           scenes.box_rays(2000, (-9, -2, -1), (9, 16, 17), seed=6), scenes.axis_rays((-9, -2, -1), (9, 16, 17), 40, seed=7)])
    _coop_same(rays, ol.oracle_closest(tris, nodes, rays), ol.oracle_any(tris, nodes, rays))


def test_coop_bumpy_negative_t_and_short_rays(bumpy_ref):
    tris, nodes, _ = bumpy_ref
    ol.emu_build(tris, nodes)
    rays = scenes.shell_rays(2500, 10.0, seed=41)
    h = ol.oracle_closest(tris, nodes, rays)
    _coop_same(rays, h)
    b = scenes.bounce_rays(rays, h, scenes.tri_normals(tris, h), seed=42)
    hb = ol.oracle_closest(tris, nodes, b)
    assert (hb["t"] < 0).mean() > 0.01
    _coop_same(b, hb, ol.oracle_any(tris, nodes, b))            # negative t ends the search at the reference's leaf
    short = b.copy()
    short["tmax"] = 3.0
    _coop_same(short, ol.oracle_closest(tris, nodes, short), ol.oracle_any(tris, nodes, short))
    inside = scenes.box_rays(1500, (-5, -5, -5), (5, 5, 5), seed=43)
    _coop_same(inside, ol.oracle_closest(tris, nodes, inside))


def test_coop_exact_ties_follow_the_reference_order(tmp_scene_dir):
    import os
    p, n, f = scenes.tie_grid(24, layers=2)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "ties.obj"), p, n, f)
    tris, nodes, _ = _reference_scene(path)
    ol.emu_build(tris, nodes)
    rays = scenes.tie_rays(24)[::5]
    _coop_same(rays, ol.oracle_closest(tris, nodes, rays), ol.oracle_any(tris, nodes, rays))


def test_coop_random_scenes_fuzz(tmp_scene_dir):
    """The fuzz scenes of test_random_scenes_fuzz (slivers, clusters, duplicated vertices, zero-area faces; leaves of
    1..64 primitives incl. blocks of more than 16 records, which take the every-lane path) in cooperative mode."""
    import os
    prod_host = __import__("conftest").load_product().host
    rng = np.random.default_rng(2025)
    for case in range(12):
        pos, faces = scenes.fuzz_scene(case, rng)
        nrm = np.tile(np.array([[0.0, 0.0, 1.0]], dtype=np.float32), (pos.shape[0], 1))
        path = scenes.write_obj(os.path.join(tmp_scene_dir, "cfuzz%d.obj" % case), pos, nrm, faces)
        max_prims = int(rng.choice([1, 2, 4, 8, 64]))
        tris, nodes, _ = prod_host.load_scene(path, max_prims)
        ol.emu_build(tris, nodes)
        lo, hi = pos.min(0) - 1.0, pos.max(0) + 1.0
        rays = np.concatenate([scenes.box_rays(600, lo, hi, seed=case), scenes.axis_rays(lo, hi, 10, seed=case),
                               scenes.pack_rays(pos[rng.integers(0, pos.shape[0], 150)], rng.normal(size=(150, 3)))])
        rays["tmax"][::5] = rng.uniform(0.1, 20.0, size=rays["tmax"][::5].shape).astype(np.float32)
        _coop_same(rays, ol.oracle_closest(tris, nodes, rays), ol.oracle_any(tris, nodes, rays))


def test_coop_degenerate_scenes(tmp_scene_dir):
    import os
    p = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float32)
    n = np.tile(np.array([[0, 0, 1]], dtype=np.float32), (4, 1))
    for name, faces, quads in (("cone.obj", [[0, 1, 2]], None), ("cquad.obj", [[0, 1, 2]], [[0, 1, 3, 2]])):
        path = scenes.write_obj(os.path.join(tmp_scene_dir, name), p, n, np.array(faces), quads)
        tris, nodes, _ = _reference_scene(path)
        ol.emu_build(tris, nodes)
        rays = scenes.box_rays(2000, (-1, -1, 0.5), (2, 2, 3), seed=45)
        _coop_same(rays, ol.oracle_closest(tris, nodes, rays))


def test_rays_with_zero_direction_components_are_culled_like_any_other_ray(bumpy_ref):
    """1/d = +-inf on an axis: the one-FMA plane evaluation gives inf - inf = NaN there and stops culling (measured on
    the GPU: one such ray walked 26 000 nodes of a 10 M-face scene). Such rays take test_wide_node_robust; results equal
    the oracle's and the walk is as short as that of the same rays tilted by a hair."""
    tris, nodes, _ = bumpy_ref
    ol.emu_build(tris, nodes)
    rng = np.random.default_rng(77)
    n = 3000
    o = rng.normal(size=(n, 3))
    o *= 12.0 / np.linalg.norm(o, axis=1, keepdims=True)          # from outside (the mesh is one-sided), towards the ball
    d = rng.uniform(-3.0, 3.0, size=(n, 3)) - o
    zero = rng.integers(0, 3, size=n)
    d[np.arange(n), zero] = 0.0                                   # one component exactly zero ...
    two = rng.random(n) < 0.3
    d[two, (zero[two] + 1) % 3] = 0.0                             # ... or two (axis-parallel)
    d[:, 0] = np.where(np.abs(d).sum(axis=1) == 0, 1.0, d[:, 0])
    d[rng.random(n) < 0.25] *= -1.0
    rays = scenes.pack_rays(o, d, 100000.0)
    want = ol.oracle_closest(tris, nodes, rays)
    assert (want["tri"] != MISS).mean() > 0.1
    st = ol.EmuStats()
    _same(ol.emu_trace(rays, stats=st), want)
    _same(ol.emu_trace(rays, schedule=9), want)
    _same(ol.emu_trace_coop(rays), want)
    _same(ol.emu_trace_coop(rays, handoff=12), want)
    _same(ol.emu_trace_coop(rays, handoff=12, resume=9), want)
    assert np.array_equal(ol.emu_trace(rays, any_hit=True) != 0, ol.oracle_any(tris, nodes, rays) != 0)
    tilted = scenes.pack_rays(o, d + 1e-4 * rng.normal(size=(n, 3)), 100000.0)
    st2 = ol.EmuStats()
    ol.emu_trace(tilted, stats=st2)
    assert st.wide_visits < 1.25 * st2.wide_visits, (st.wide_visits, st2.wide_visits)


def test_leaf_blocks_with_explicit_boxes(tmp_scene_dir):
    """Leaf blocks normally carry no box: the kernels recompute the min / max of the record's vertices, which is what the
    reference's builder stores (CLBVHnode.cpp:18-23). Blocks of several records, and leaves of a caller-supplied tree whose
    box is something else, keep an explicit box -- here leaf boxes SHRUNK below their triangles (a valid tree: children
    stay inside parents), so a walk that used the vertices' box instead would accept hits the reference culls."""
    import os
    p, n, f = scenes.displaced_sphere(4)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "boxed.obj"), p, n, f)
    rays = scenes.shell_rays(30000, 10.0, seed=91)
    for max_prims in (4, 16):
        tris, nodes, _ = _reference_scene(path, max_prims)
        st = ol.emu_build(tris, nodes)
        words_plain = st.leaf_words
        _same(ol.emu_trace(rays), ol.oracle_closest(tris, nodes, rays))
        nodes = nodes.copy()
        raw = nodes.view(np.uint8).reshape(-1, 48)
        box = raw[:, :32].view(np.float32).reshape(-1, 8)          # bmin.xyzw, bmax.xyzw
        nprim = raw[:, 36:38].view(np.uint16).reshape(-1)
        leaf = nprim > 0
        centre = 0.5 * (box[leaf, 0:3] + box[leaf, 4:7])
        box[leaf, 0:3] = centre + 0.35 * (box[leaf, 0:3] - centre)
        box[leaf, 4:7] = centre + 0.35 * (box[leaf, 4:7] - centre)
        st = ol.emu_build(tris, nodes)
        assert st.leaf_words > words_plain + 1.8 * leaf.sum()      # (nearly) every block went without a box before and carries two box words now
        want = ol.oracle_closest(tris, nodes, rays)
        assert (want["tri"] != MISS).mean() < 0.8 * (ol.oracle_closest(*_reference_scene(path, max_prims)[:2], rays)["tri"] != MISS).mean()
        _same(ol.emu_trace(rays), want)
        _same(ol.emu_trace(rays, schedule=5), want)
        _same(ol.emu_trace_coop(rays[:6000], handoff=5), want[:6000])
        _same(ol.emu_trace_coop(rays[:6000]), want[:6000])


def test_scaled_and_offset_scenes(tmp_scene_dir):
    """The packed-fp16 node test works in a per-visit frame (entry distance, power-of-two scale): tiny scenes, huge scenes
    and scenes far from the origin (coordinates ~1e5 with metre-sized triangles, where fp32 itself leaves only ~7 bits
    inside a leaf) must still give the oracle's hits, with tmax left at its default and tightened to scene size."""
    import os
    p, n, f = scenes.displaced_sphere(4)
    for k, (scale, centre) in enumerate(((1.0e-4, (0.0, 0.0, 0.0)), (1.0e4, (0.0, 0.0, 0.0)), (1.0, (1.0e5, -2.0e5, 5.0e4)),
                                         (1.0e-3, (-300.0, 700.0, 90.0)), (3.0e5, (1.0e7, 1.0e7, -1.0e7)))):
        q = (p.astype(np.float64) * scale + np.asarray(centre)).astype(np.float32)
        path = scenes.write_obj(os.path.join(tmp_scene_dir, "scaled%d.obj" % k), q, n, f)
        tris, nodes, _ = _reference_scene(path)
        ol.emu_build(tris, nodes)
        radius = 10.0 * scale
        for tmax in (max(100000.0, 40.0 * radius), 3.5 * radius):      # the default (or beyond the scene), and one that cuts rays short
            rays = np.concatenate([scenes.shell_rays(6000, radius, seed=61 + k, centre=centre, tmax=tmax),
                                   scenes.box_rays(3000, np.asarray(centre) - 1.5 * radius, np.asarray(centre) + 1.5 * radius, seed=71 + k, tmax=tmax)])
            want = ol.oracle_closest(tris, nodes, rays)
            if scale >= 1.0:                                      # (tiny triangles fall under the reference's det < 1e-8 cull: all misses)
                assert (want["tri"] != MISS).mean() > 0.25
            _same(ol.emu_trace(rays), want)
            _same(ol.emu_trace(rays, schedule=3), want)
            _same(ol.emu_trace_coop(rays[:3000], handoff=7), want[:3000])
            assert np.array_equal(ol.emu_trace(rays, any_hit=True) != 0, ol.oracle_any(tris, nodes, rays) != 0)


def test_fast_node_test_never_culls_what_the_exact_one_lets_through():
    """Runs last in this file: every emulated walk above (all scenes, schedules and the cooperative mode's solo prefixes)
    compared, node by node, the fast wide-node test with test_wide_node_robust on the same node, ray and `best`
    (B2_EMU_CHECK_CULLING): the fast test may pass more children, never fewer."""
    assert ol.emu_culling_violations() == 0
