"""Developer script (GPU): the per-bounce traversal tails of a rank's share of a tiled 4K frame, with the cooperative tail
mode off / on. usage: dev_tail.py [faces=10000000] [world=8]"""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
faces = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
W, H = 3840, 2160
if faces > 0:
    path = "/tmp/b2rt_scenes/scatter_%d.obj" % faces
    t0 = time.time()
    if not os.path.exists(path + ".done"):
        subprocess.check_call([os.path.join(os.path.dirname(prod.lib_path()), "scenegen"), "scattered", path, str(faces), "50.0", "0.05", "0.5", "11"], stdout=subprocess.DEVNULL)
        open(path + ".done", "w").close()
    print("scene written %.1f s" % (time.time() - t0), flush=True)
    cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
else:
    import scenes
    path = scenes.CORNELL
    cam = {}
t0 = time.time()
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
print("loaded + built %.1f s: %d tris %d nodes" % (time.time() - t0, t.shape[0], n.shape[0]), flush=True)
with prod.Context(0) as ctx:
    t0 = time.time()
    ctx.upload_scene(t, n, m)
    print("uploaded %.1f s" % (time.time() - t0), ctx.scene_info(), flush=True)
    ctx.resize(W, H)
    ctx.set_option(cap.OPT_RENDER_MODE, 0)
    plan = prod.sharding.BandPlan(W, H, world, band_rows=8)

    def frames(fn, k=6):
        for f in (1, 2):
            ctx.set_frame(f, 4, **cam); fn()
        ctx.finish()
        t0 = time.perf_counter()
        for f in range(3, 3 + k):
            ctx.set_frame(f, 4, **cam); fn()
        ctx.finish()
        return (time.perf_counter() - t0) / k * 1e3

    ref = None
    for coop in (0, 8, 16):
        for lanes in (1, 2):
            ctx.set_option(cap.OPT_COOP_MAX, coop)
            ctx.set_option(cap.OPT_WAVEFRONT_LANES, lanes)
            share = frames(lambda: plan.render(ctx, 0))
            print("coop %2d lanes %d: 1/%d share %.3f ms" % (coop, lanes, world, share), flush=True)
    for coop in (0, 8):
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_WAVEFRONT_LANES, 0)
        full = frames(lambda: ctx.execute(W * H), 3)
        img = ctx.read_pixels().copy()
        same = True if ref is None else bool(np.array_equal(ref.view(np.uint32), img.view(np.uint32)))
        ref = img
        print("coop %2d: full frame %.3f ms, identical to previous: %s" % (coop, full, same), flush=True)
    ctx.set_option(cap.OPT_RENDER_MODE, 1)
    print("megakernel: 1/%d share %.3f ms, full %.3f ms" % (world, frames(lambda: plan.render(ctx, 0)), frames(lambda: ctx.execute(W * H), 3)), flush=True)
    ctx.set_option(cap.OPT_RENDER_MODE, 0)
    # coarser bands: a rank's rays touch a smaller part of the BVH (L2 / DRAM locality) at the price of load balance
    ctx.set_option(cap.OPT_COOP_MAX, 8)
    ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
    for rows in (8, 30, 90, 270):
        p2 = prod.sharding.BandPlan(W, H, world, band_rows=rows)
        ts = [frames(lambda: p2.render(ctx, r), 4) for r in (0, world // 2, world - 1)]
        print("band_rows %3d: 1/%d share of ranks 0 / %d / %d: %s ms" % (rows, world, world // 2, world - 1, " ".join("%.3f" % t for t in ts)), flush=True)
    # per-stage device times of the share (events after every stage of the launch's one wavefront)
    for coop in (0, 8):
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
        ctx.set_option(cap.OPT_STAGE_TIMES, 1)
        for f in (20, 21, 22):
            ctx.set_frame(f, 4, **cam); plan.render(ctx, 0); ctx.finish()
        st = ctx.stage_times()
        ctx.set_option(cap.OPT_STAGE_TIMES, 0)
        print("coop %d stages (ms): %s  | sum %.3f" % (coop, " ".join("%s %.3f" % (k, v) for k, v in st), sum(v for _, v in st)), flush=True)
    for coop in (0, 8):
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
        ctx.set_option(cap.OPT_STAGE_TIMES, 1)
        for f in (30, 31, 32):
            ctx.set_frame(f, 4, **cam); ctx.execute(W * H); ctx.finish()
        st = ctx.stage_times()
        ctx.set_option(cap.OPT_STAGE_TIMES, 0)
        print("FULL frame coop %d stages (ms): %s  | sum %.3f" % (coop, " ".join("%s %.3f" % (k, v) for k, v in st), sum(v for _, v in st)), flush=True)
    # counted share: how much work goes through the tail kernel, and the worst solo ray
    for coop in (0, 8):
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
        ctx.set_option(cap.OPT_COUNTERS, 1)
        ctx.reset_counters()
        ctx.set_frame(9, 4, **cam); plan.render(ctx, 0); ctx.finish()
        c = ctx.counters()
        ctx.set_option(cap.OPT_COUNTERS, 0)
        print("coop %d counted share:" % coop, {k: c[k] for k in ("rays", "wide_nodes", "leaf_blocks", "max_steps_per_ray", "coop_rays", "coop_steps", "stack_overflows")}, flush=True)
