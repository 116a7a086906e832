"""-m gpu, round 2: the cooperative tail mode, direct comparisons with the VERBATIM reference build (oracle/_ref),
the advisor's regression cases, and a >= 1 M-face scattered scene checked against oracle-rendered pixels."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
import scenes
from test_gpu_parity import _check_hits, _render

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF


def _counted(product, ctx, fn):
    ctx.set_option(product.capi.OPT_COUNTERS, 1)
    ctx.reset_counters()
    try:
        out = fn()
        return out, ctx.counters()
    finally:
        ctx.set_option(product.capi.OPT_COUNTERS, 0)


def test_cooperative_tail_is_result_neutral(product, bumpy_ref, tmp_scene_dir):
    """B2RT_OPT_COOP_MAX: a dry warp hands its last rays to trace_tail_kernel (32 lanes per ray, csrc/coop.cuh). Hits must
    not depend on the threshold -- primary rays, negative-t bounce rays, exact cross-leaf ties, any-hit -- and the
    counting build must show that the tail kernel really ran."""
    cap = product.capi
    tris, nodes, mats = bumpy_ref
    rays = scenes.shell_rays(150000, 10.0, seed=151)
    want = ol.oracle_closest(tris, nodes, rays)
    b = scenes.bounce_rays(rays, want, scenes.tri_normals(tris, want), seed=152)
    wb = ol.oracle_closest(tris, nodes, b)
    assert (wb["t"] < 0).mean() > 0.01
    occ_want = ol.oracle_any(tris, nodes, b)
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        for coop in (0, 1, 8, 16):
            ctx.set_option(cap.OPT_COOP_MAX, coop)
            got, c = _counted(product, ctx, lambda: ctx.trace_closest(rays))
            _check_hits(got, want)
            assert c["rays"] == rays.shape[0] and c["stack_overflows"] == 0
            assert (c["coop_rays"] > 0) == (coop > 0)
            _check_hits(ctx.trace_closest(b), wb)
            assert np.array_equal(ctx.trace_any(b) != 0, occ_want != 0)
            for n in (1, 31, 33, 1000):                           # launches so small that (almost) every ray ends in the tail kernel
                _check_hits(ctx.trace_closest(b[:n]), wb[:n])
        # launches of 16 rays: the pool is dry at once and the warp is under the threshold from the start, so rays go to the tail
        # kernel as soon as every live lane has work stacked up (lanes about to finish are waited for)
        ctx.set_option(cap.OPT_COOP_MAX, 16)
        got, c = _counted(product, ctx, lambda: np.concatenate([ctx.trace_closest(b[i:i + 16]) for i in range(0, 4000, 16)]))
        _check_hits(got, wb[:4000])
        assert c["coop_rays"] > 100
        with pytest.raises(product.B2RTError):
            ctx.set_option(cap.OPT_COOP_MAX, 17)
    p, n, f = scenes.tie_grid(24, layers=2)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "ties_coop.obj"), p, n, f)
    t2, n2, m2 = product.host.load_scene(path, 4)
    tr = scenes.tie_rays(24)
    tw = ol.oracle_closest(t2, n2, tr)
    with product.Context(0) as ctx:
        ctx.upload_scene(t2, n2, m2)
        for coop in (0, 16):
            ctx.set_option(cap.OPT_COOP_MAX, coop)
            _check_hits(ctx.trace_closest(tr), tw)
            _check_hits(np.concatenate([ctx.trace_closest(tr[i:i + 16]) for i in range(0, 1600, 16)]), tw[:1600])


def test_two_step_tail_is_result_neutral(product, bumpy_ref, cornell_ref):
    """B2RT_OPT_RESUME_MAX: dry warps suspend their last rays early, a second persistent launch re-packs the suspended rays
    and walks on (Lane::resume), the cooperative kernel finishes what that pass leaves. Hits and frames must not depend on
    either threshold, and the counting build must show that rays really were resumed."""
    cap = product.capi
    tris, nodes, mats = bumpy_ref
    rays = scenes.shell_rays(120000, 10.0, seed=171)
    want = ol.oracle_closest(tris, nodes, rays)
    b = scenes.bounce_rays(rays, want, scenes.tri_normals(tris, want), seed=172)
    wb = ol.oracle_closest(tris, nodes, b)
    occ_want = ol.oracle_any(tris, nodes, b)
    W, H = 301, 203
    cam = dict(pos=(0.0, -14.0, 2.0), front=(0.0, 1.0, -0.1), up=(0.0, 0.0, 1.0))
    frames = []
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        ctx.resize(W, H)
        ctx.set_option(cap.OPT_RENDER_MODE, 0)
        for coop, resume in ((8, 0), (8, 16), (1, 16), (4, 9), (2, 3)):
            ctx.set_option(cap.OPT_COOP_MAX, coop)
            ctx.set_option(cap.OPT_RESUME_MAX, resume)
            got, c = _counted(product, ctx, lambda: ctx.trace_closest(rays))
            _check_hits(got, want)
            assert c["rays"] == rays.shape[0] and c["stack_overflows"] == 0
            assert (c["resumed_rays"] > 0) == (resume > coop), (coop, resume, c)
            _check_hits(ctx.trace_closest(b), wb)
            assert np.array_equal(ctx.trace_any(b) != 0, occ_want != 0)
            for n in (1, 31, 33, 1000):
                _check_hits(ctx.trace_closest(b[:n]), wb[:n])
            got = np.concatenate([ctx.trace_closest(b[i:i + 16]) for i in range(0, 1600, 16)])     # every ray is suspended at once
            _check_hits(got, wb[:1600])
            for fc in (0, 1, 2):                                   # frameCount 0 overwrites the image, 1 and 2 accumulate
                ctx.set_frame(fc, 4, **cam)
                ctx.execute(W * H)
            frames.append(ctx.read_pixels().copy())
        with pytest.raises(product.B2RTError):
            ctx.set_option(cap.OPT_RESUME_MAX, 17)
    for f in frames[1:]:
        assert np.array_equal(f.view(np.uint32), frames[0].view(np.uint32))


def test_cuda_path_against_the_verbatim_reference_build(product, bumpy_ref):
    """no port in between -- CUDA hits and a CUDA frame compared directly with the reference's own
    Intersect() / KernelEntry compiled from /root/reference (oracle/_ref)."""
    if ol.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not available on this machine")
    tris, nodes, mats = bumpy_ref
    rays = scenes.shell_rays(200000, 10.0, seed=161)
    ref = ol.ref_closest(tris, nodes, rays)
    hit = ref["hit"] != 0
    b = scenes.bounce_rays(rays, ol.oracle_closest(tris, nodes, rays), scenes.tri_normals(tris, ol.oracle_closest(tris, nodes, rays)), seed=162)
    refb = ol.ref_closest(tris, nodes, b)
    hitb = refb["hit"] != 0
    W, H = 256, 192
    cam = dict(pos=(0.0, -30.0, 4.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in (1, 2, 3):
        ol.ref_render(tris, nodes, mats, want, W, H, fc, 4, **cam)
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        for r, want_hits, m in ((rays, ref, hit), (b, refb, hitb)):
            got = ctx.trace_closest(r)
            assert np.array_equal(got["tri"] != MISS, m)
            assert np.array_equal(got["tri"][m], want_hits["tri"][m].astype(np.uint32))
            assert np.array_equal(got["t"][m].view(np.uint32), want_hits["t"][m].view(np.uint32))
        assert (refb["t"][hitb] < 0).mean() > 0.01                 # the negative-t trap is in the sample
        img = _render(ctx, W, H, (1, 2, 3), 4, **cam)
    assert scenes.psnr(img[:, :3], want[:, :3]) >= 50.0
    assert (img[:, :3] == want[:, :3]).all(axis=1).mean() > 0.95


def test_oracle_and_host_suites_on_this_machine():
    """the pins port <-> verbatim reference <-> golden files and host arrays <-> reference builder are CPU
    tests; run them on the GPU box too, so that the chain GPU = port = reference is closed on one machine."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", os.path.join(root, "tests", "test_oracle.py"),
                        os.path.join(root, "tests", "test_host_scene.py")], cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_wavefront_lane_split_keeps_every_work_item(product, cornell_ref):
    """n = 1 (mod 64) work items split over 2 or 3 wavefront lanes used to lose the last item."""
    cap = product.capi
    tris, nodes, mats = cornell_ref
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        for W, H, lanes in ((65, 1, 2), (97, 1, 3), (1025, 1025, 0), (1025, 1025, 3)):
            frames = {}
            for mode in (1, 0):
                ctx.set_option(cap.OPT_RENDER_MODE, mode)
                ctx.set_option(cap.OPT_WAVEFRONT_LANES, lanes)
                ctx.resize(W, H)
                ctx.set_frame(1, 3)
                ctx.execute(W * H)
                frames[mode] = ctx.read_pixels().copy()
            assert np.array_equal(frames[0].view(np.uint32), frames[1].view(np.uint32)), (W, H, lanes)
            assert frames[0][-1, :3].max() > 0.0


def test_launches_on_different_streams_may_overlap(product, bumpy_ref):
    """every launch has its own ray counter, so traces enqueued on different caller streams (and on the
    context's stream) at the same time do not disturb each other."""
    import torch
    tris, nodes, mats = bumpy_ref
    sets = [scenes.shell_rays(400000, 10.0, seed=171 + i) for i in range(3)]
    wants = [ol.oracle_closest(tris, nodes, r) for r in sets]
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        d_rays = [torch.from_numpy(r.view(np.float32).reshape(-1, 8)).to("cuda:0") for r in sets]
        d_hits = [torch.empty((r.shape[0], 4), dtype=torch.float32, device="cuda:0") for r in sets]
        streams = [torch.cuda.Stream(device="cuda:0") for _ in range(2)]
        torch.cuda.synchronize()
        for rep in range(4):
            for h in d_hits:
                h.zero_()
            torch.cuda.synchronize()
            ctx.trace_closest_device(d_rays[0].data_ptr(), sets[0].shape[0], d_hits[0].data_ptr(), streams[0].cuda_stream)
            ctx.trace_closest_device(d_rays[1].data_ptr(), sets[1].shape[0], d_hits[1].data_ptr(), streams[1].cuda_stream)
            ctx.trace_closest_device(d_rays[2].data_ptr(), sets[2].shape[0], d_hits[2].data_ptr(), 0)   # the context's own stream
            torch.cuda.synchronize()
            ctx.finish()
            for h, w in zip(d_hits, wants):
                _check_hits(h.cpu().numpy().view(product.HIT_DTYPE).reshape(-1), w)


def test_million_face_scattered_scene_against_oracle_pixels(product, tmp_scene_dir):
    """BASELINE.json configs[3] at >= 1 M OBJ faces (2 M CLTriangle): pixels of the 4-bounce frame, rendered
    through the wavefront path with the tail kernel, against the SAME pixels rendered by the oracle (oracle_render over gid
    ranges), and incoherent rays against the oracle's hits."""
    path = os.path.join(tmp_scene_dir, "scatter_1m.obj")
    assert product.host.write_scattered_obj(path, 1000000, extent=50.0, edge_min=0.25, edge_max=1.0, seed=11) == 1000000
    tris, nodes, mats = product.host.load_scene(path, 4)
    assert tris.shape[0] == 2000000
    W, H, bounces = 1280, 720, 4
    cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
    rows = (5, 200, 359, 360, 511, 700)                          # image rows rendered by the oracle
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in (1, 2):
        for y in rows:
            ol.oracle_render(tris, nodes, mats, want, W, H, fc, bounces, gid0=y * W, gid1=(y + 1) * W, **cam)
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        img = _render(ctx, W, H, (1, 2), bounces, **cam)
        sel = np.concatenate([np.arange(y * W, (y + 1) * W) for y in rows])
        assert scenes.psnr(img[sel, :3], want[sel, :3]) >= 50.0
        assert (img[sel, :3] == want[sel, :3]).all(axis=1).mean() > 0.95
        assert 0.2 < (img[sel, :3].max(axis=1) > 0).mean()
        rays = scenes.box_rays(300000, (-60, -60, -60), (60, 60, 60), seed=181)
        got = ctx.trace_closest(rays)
        _check_hits(got, ol.oracle_closest(tris, nodes, rays))
        assert 0.2 < (got["tri"] != MISS).mean() < 0.999
        assert np.array_equal(ctx.trace_any(rays) != 0, got["tri"] != MISS)


# ---- multi-GPU behind the C ABI (csrc/multi.cu) ---------------------------------------------------------------------
def _device_list(product, want):
    """`want` distinct devices when the box has them; else the same device several times (B2RT_ALLOW_DUPLICATE_DEVICES:
    every mention is a member with its own streams and buffers -- partition, worker threads and store-through run the
    same code, only the NVLink hop is missing)."""
    n = product.device_count()
    if n >= want:
        return list(range(want))
    os.environ["B2RT_ALLOW_DUPLICATE_DEVICES"] = "1"
    return [0] * want


def test_device_group_frame_and_streams(product, bumpy_ref, cornell_ref):
    """b2rt_create_multi: one handle, several devices. Frames (progressive accumulation over several frames, whole and
    clipped bands, both render modes) must equal the single-device frames bit for bit; host ray streams likewise."""
    cap = product.capi
    for ref, cam, W, H in ((cornell_ref, {}, 203, 131), (bumpy_ref, dict(pos=(0.0, -30.0, 4.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0)), 320, 67)):
        tris, nodes, mats = ref
        with product.Context(0) as one:
            one.upload_scene(tris, nodes, mats)
            alone = _render(one, W, H, (1, 2, 3), 4, **cam)
        for n_dev in (2, 3):
            with product.Context(_device_list(product, n_dev)) as grp:
                assert grp.group_size() == n_dev
                grp.upload_scene(tris, nodes, mats)
                for mode in (0, 1, 2):
                    grp.set_option(cap.OPT_RENDER_MODE, mode)
                    img = _render(grp, W, H, (1, 2, 3), 4, **cam)
                    assert np.array_equal(img.view(np.uint32), alone.view(np.uint32)), (n_dev, mode)
                info = grp.scene_info()
                assert info["n_triangles"] == tris.shape[0]
    tris, nodes, mats = bumpy_ref
    rays = scenes.shell_rays(300001, 10.0, seed=191)
    want = ol.oracle_closest(tris, nodes, rays)
    with product.Context(_device_list(product, 2)) as grp:
        grp.upload_scene(tris, nodes, mats)
        _check_hits(grp.trace_closest(rays), want)
        assert np.array_equal(grp.trace_any(rays) != 0, want["tri"] != MISS)
        _check_hits(grp.trace_closest(rays[:5]), want[:5])
        with pytest.raises(product.B2RTError):
            grp.execute_bands(0, 64, 128, 2)                        # addresses one device


def test_device_group_orders_peer_stores_behind_the_roots_reads(product, cornell_ref):
    """b2rt_read_pixels is asynchronous (clEnqueueReadBuffer, non-blocking): a frame enqueued right behind it must not
    change what the read returns, although the peers' kernels write the root's image from their own streams."""
    tris, nodes, mats = cornell_ref
    W, H, frames = 1024, 768, (1, 2, 3, 4, 5)
    with product.Context(0) as one:
        one.upload_scene(tris, nodes, mats)
        one.resize(W, H)
        want = []
        for fc in frames:
            one.set_frame(fc, 4)
            one.execute(W * H)
            want.append(one.read_pixels().copy())
    with product.Context(_device_list(product, 3)) as grp:
        grp.upload_scene(tris, nodes, mats)
        grp.resize(W, H)
        got = [np.zeros((W * H, 4), dtype=np.float32) for _ in frames]
        for g in got:
            grp.host_register(g)
        try:
            for rep in range(2):                                # frame 1 ignores the image's old content; the second pass runs with the render-mode choice settled
                for fc, g in zip(frames, got):
                    grp.set_frame(fc, 4)
                    grp.execute(W * H)
                    grp.read_pixels_async(g)                    # no finish: the next frame is enqueued behind the copy
                grp.finish()
                if rep == 0:
                    continue
                for fc, g, w in zip(frames, got, want):
                    assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), fc
        finally:
            grp.finish()
            for g in got:
                grp.host_unregister(g)


def test_cpp_program_drives_a_device_group(product, tmp_scene_dir, cornell_ref):
    """a C++ program written against CLRaytracer renders a multi-GPU frame with no Python in the path."""
    from test_cpp_host import _build
    exe = _build(tmp_scene_dir)
    W, H, frames, bounces = 160, 120, 5, 4
    outs = []
    devs = _device_list(product, 2)
    for tag, extra in (("one", []), ("two", [",".join(str(d) for d in devs)])):
        out = os.path.join(tmp_scene_dir, "cornell_%s.raw" % tag)
        r = subprocess.run([exe, scenes.CORNELL, str(W), str(H), str(frames), str(bounces), out] + extra, capture_output=True, text=True,
                           env=dict(os.environ, B2RT_ALLOW_DUPLICATE_DEVICES="1"))
        assert r.returncode == 0, r.stderr
        assert ("%d device(s)" % (2 if extra else 1)) in r.stdout
        outs.append(np.fromfile(out, dtype=np.float32).reshape(-1, 4))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    tris, nodes, mats = cornell_ref
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in range(1, frames + 1):
        ol.oracle_render(tris, nodes, mats, want, W, H, fc, bounces)
    assert scenes.psnr(outs[1][:, :3], want[:, :3]) >= 50.0


# ---- refit (csrc/refit.cu) ------------------------------------------------------------------------------------------------
def _deform(tris, scale, wobble):
    """New vertex positions as a function of the OLD position only: the loader's duplicate copies of a face stay
    bit-identical rotations of each other. Normals are recomputed per face (flat) to change the shading records too."""
    t = tris.copy()
    f = t.view(np.float32).reshape(-1, 64)
    for v in (0, 20, 40):                                       # v1 / v2 / v3: position at float 0, normal at float 8 of each 80-byte vertex
        p = f[:, v:v + 3]
        q = p * np.float32(scale)
        q[:, 2] += np.float32(wobble) * np.sin(p[:, 0] * np.float32(0.7)).astype(np.float32)
        f[:, v:v + 3] = q
    e1, e2 = f[:, 20:23] - f[:, 0:3], f[:, 40:43] - f[:, 0:3]
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    for v in (0, 20, 40):
        f[:, v + 8:v + 11] = nrm.astype(np.float32)
    return t


def test_refit_scene_matches_the_oracle_on_the_refitted_tree(product, bumpy_ref):
    """b2rt_refit_scene: deformed vertices, same topology. (1) refitting with UNCHANGED data reproduces the host builder's
    boxes bit for bit; (2) after a deformation the CUDA hits and frames equal the oracle's walk over (new triangles,
    refitted nodes read back from the device); (3) data that breaks a duplicate pair is refused."""
    tris, nodes, mats = bumpy_ref
    rays = scenes.shell_rays(200000, 12.0, seed=201)
    with product.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        before = ctx.trace_closest(rays)
        ctx.refit_scene(tris)
        assert np.array_equal(ctx.read_nodes(), nodes)                      # boxes, offsets, counts, axes: all as the host built them
        assert np.array_equal(ctx.trace_closest(rays).view(np.uint32), before.view(np.uint32))
        for scale, wobble in ((1.2, 0.8), (0.6, 0.0), (1.0, 2.5)):
            new = _deform(tris, scale, wobble)
            ctx.refit_scene(new)
            rn = ctx.read_nodes()
            assert np.array_equal(rn[:, 32:39], nodes[:, 32:39])             # topology untouched
            want = ol.oracle_closest(new, rn, rays)
            got = ctx.trace_closest(rays)
            _check_hits(got, want)
            assert 0.05 < (got["tri"] != MISS).mean()
            b = scenes.bounce_rays(rays, want, scenes.tri_normals(new, want), seed=202)
            _check_hits(ctx.trace_closest(b), ol.oracle_closest(new, rn, b))
            assert np.array_equal(ctx.trace_any(b) != 0, ol.oracle_any(new, rn, b) != 0)
            W, H = 160, 120
            cam = dict(pos=(0.0, -36.0, 4.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
            ref_img = np.zeros((W * H, 4), dtype=np.float32)
            for fc in (1, 2):
                ol.oracle_render(new, rn, mats, ref_img, W, H, fc, 3, **cam)
            assert scenes.psnr(_render(ctx, W, H, (1, 2), 3, **cam)[:, :3], ref_img[:, :3]) >= 50.0      # the shading records follow too
            ctx.set_option(product.capi.OPT_TRAVERSAL, 1)                # the reference-layout walk sees the same refitted arrays
            _check_hits(ctx.trace_closest(rays), want)
            ctx.set_option(product.capi.OPT_TRAVERSAL, 0)
        broken = tris.copy()
        broken.view(np.float32).reshape(-1, 64)[0, 0] += np.float32(0.25)      # one copy of a duplicated face moves alone
        with pytest.raises(product.B2RTError) as e:
            ctx.refit_scene(broken)
        assert e.value.status == -50                                          # CL_INVALID_ARG_VALUE
        with pytest.raises(product.B2RTError):
            ctx.refit_scene(tris[:-2])
        ctx.refit_scene(tris)                                                 # and back
        assert np.array_equal(ctx.trace_closest(rays).view(np.uint32), before.view(np.uint32))


def test_leaf_blocks_with_explicit_boxes_on_the_gpu(product, tmp_scene_dir):
    """The compact leaf block (no stored box: min / max of the record's vertices) against blocks that must keep their box:
    several records per leaf (max_prims 16) and leaf boxes shrunk below their triangles. Same check as the CPU emulation's
    (tests/test_emu_traversal.py), through the C ABI, solo and cooperative walks, plus a refit of the compact scene."""
    cap = product.capi
    p, n, f = scenes.displaced_sphere(4)
    path = scenes.write_obj(os.path.join(tmp_scene_dir, "boxed_gpu.obj"), p, n, f)
    rays = scenes.shell_rays(60000, 10.0, seed=191)
    for max_prims in (4, 16):
        tris, nodes, mats = product.host.load_scene(path, max_prims)
        shrunk = nodes.copy()
        raw = shrunk.view(np.uint8).reshape(-1, 48)
        box = raw[:, :32].view(np.float32).reshape(-1, 8)
        leaf = raw[:, 36:38].view(np.uint16).reshape(-1) > 0
        centre = 0.5 * (box[leaf, 0:3] + box[leaf, 4:7])
        box[leaf, 0:3] = centre + 0.35 * (box[leaf, 0:3] - centre)
        box[leaf, 4:7] = centre + 0.35 * (box[leaf, 4:7] - centre)
        for nd in (nodes, shrunk):
            want = ol.oracle_closest(tris, nd, rays)
            with product.Context(0) as ctx:
                ctx.upload_scene(tris, nd, mats)
                for coop in (0, 16):
                    ctx.set_option(cap.OPT_COOP_MAX, coop)
                    _check_hits(ctx.trace_closest(rays), want)
                    _check_hits(np.concatenate([ctx.trace_closest(rays[i:i + 16]) for i in range(0, 800, 16)]), want[:800])
                assert np.array_equal(ctx.trace_any(rays) != 0, ol.oracle_any(tris, nd, rays) != 0)


def test_scaled_and_offset_scenes_on_the_gpu(product, tmp_scene_dir):
    """The packed-fp16 wide-node test evaluates slabs in a per-visit frame (relative to the entry distance, scaled by a
    power of two): scenes of very different size and scenes far from the origin must give the oracle's hits bit for bit
    (the same cases run through the CPU emulation with the node-by-node culling check, tests/test_emu_traversal.py)."""
    cap = product.capi
    p, n, f = scenes.displaced_sphere(4)
    for k, (scale, centre) in enumerate(((1.0e4, (0.0, 0.0, 0.0)), (1.0, (1.0e5, -2.0e5, 5.0e4)), (1.0e-3, (-300.0, 700.0, 90.0)),
                                         (3.0e5, (1.0e7, 1.0e7, -1.0e7)))):
        q = (p.astype(np.float64) * scale + np.asarray(centre)).astype(np.float32)
        path = scenes.write_obj(os.path.join(tmp_scene_dir, "scaled_gpu%d.obj" % k), q, n, f)
        tris, nodes, mats = product.host.load_scene(path, 4)
        radius = 10.0 * scale
        with product.Context(0) as ctx:
            ctx.upload_scene(tris, nodes, mats)
            for tmax in (max(100000.0, 40.0 * radius), 3.5 * radius):
                rays = np.concatenate([scenes.shell_rays(40000, radius, seed=61 + k, centre=centre, tmax=tmax),
                                       scenes.box_rays(20000, np.asarray(centre) - 1.5 * radius, np.asarray(centre) + 1.5 * radius, seed=71 + k, tmax=tmax)])
                want = ol.oracle_closest(tris, nodes, rays)
                for coop in (8, 0):
                    ctx.set_option(cap.OPT_COOP_MAX, coop)
                    _check_hits(ctx.trace_closest(rays), want)
                assert np.array_equal(ctx.trace_any(rays) != 0, ol.oracle_any(tris, nodes, rays) != 0)
