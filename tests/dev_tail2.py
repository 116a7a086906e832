"""Developer script (GPU): a rank's 1/world share (and the full frame) of a tiled 4K frame under wavefront settings:
wavefronts in flight x grid split x cooperative-tail threshold x persistent grid size.
usage: dev_tail2.py [faces=10000000] [world=8]   (faces=0: cornell.obj)"""
import os, sys, time, subprocess, itertools
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
    if not os.path.exists(path + ".done"):
        subprocess.check_call([os.path.join(os.path.dirname(prod.host.lib_path()), "scenegen"), "scattered", path, str(faces), "50.0", "0.05", "0.5", "11"], stdout=subprocess.DEVNULL)
        open(path + ".done", "w").close()
    cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
else:
    import scenes
    path, cam = scenes.CORNELL, {}
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
with prod.Context(0) as ctx:
    ctx.upload_scene(t, n, m)
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

    def setopts(lanes, split, coop, bps):
        ctx.set_option(cap.OPT_WAVEFRONT_LANES, lanes)
        ctx.set_option(cap.OPT_WAVEFRONT_GRID_SPLIT, split)
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_BLOCKS_PER_SM, bps)

    print("== every rank's share, defaults, two passes", flush=True)
    setopts(1, 0, 8, 0)
    for rep in range(2):
        print("   " + " ".join("%.3f" % frames(lambda: plan.render(ctx, r), 4) for r in range(world)), flush=True)
    print("== share of rank 0: lanes, split, coop, blocks/SM", flush=True)
    for lanes, split, coop, bps in [] if os.environ.get("QUICK") else [(1, 0, 8, 0), (1, 0, 4, 0), (1, 0, 2, 0), (1, 0, 8, 4), (2, 0, 8, 0), (2, 1, 8, 0), (2, 1, 4, 0), (3, 1, 8, 0), (4, 1, 8, 0), (4, 1, 4, 0),
                                    (4, 0, 8, 0), (2, 1, 12, 0), (4, 1, 12, 0)]:
        setopts(lanes, split, coop, bps)
        print("   lanes %d split %d coop %2d bps %d: %.3f ms" % (lanes, split, coop, bps, frames(lambda: plan.render(ctx, 0))), flush=True)
    print("== full frame", flush=True)
    for lanes, split, coop, bps in [(0, 0, 8, 0)] if os.environ.get("QUICK") else [(0, 0, 8, 0), (2, 1, 8, 0), (4, 1, 8, 0), (4, 0, 8, 0), (4, 1, 4, 0)]:
        setopts(lanes, split, coop, bps)
        print("   lanes %d split %d coop %2d bps %d: %.3f ms" % (lanes, split, coop, bps, frames(lambda: ctx.execute(W * H), 3)), flush=True)
