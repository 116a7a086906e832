"""Developer script (GPU): one rank's share of the tiled 4K frame on the 10 M-triangle scene at several persistent grid sizes
(B2RT_OPT_BLOCKS_PER_SM) and cooperative thresholds. usage: dev_share_grid.py [faces=10000000] [world=8]"""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
faces = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
W, H = 3840, 2160
path = "/tmp/b2rt_scenes/scatter_%d.obj" % faces
if not os.path.exists(path):
    subprocess.check_call([os.path.join(os.path.dirname(prod.lib_path()), "scenegen"), "scattered", path, str(faces), "50.0", "0.05", "0.5", "11"], stdout=subprocess.DEVNULL)
cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
with prod.Context(0) as ctx:
    ctx.upload_scene(t, n, m)
    ctx.resize(W, H)
    ctx.set_option(cap.OPT_RENDER_MODE, 0)
    ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
    plan = prod.sharding.BandPlan(W, H, world, band_rows=8)

    def run(frames, f0):
        for f in range(f0, f0 + frames):
            ctx.set_frame(f, 4, **cam)
            if world > 1:
                plan.render(ctx, 0)
            else:
                ctx.execute(W * H)
            ctx.finish()
    for rep in range(2):
        for bps, refill in ((0, 6), (8, 6), (7, 6), (6, 6), (5, 6), (4, 6), (0, 3), (0, 10), (6, 10)):
            ctx.set_option(cap.OPT_BLOCKS_PER_SM, bps)
            ctx.set_option(cap.OPT_REFILL_MIN, refill)
            run(3, 1)
            t0 = time.perf_counter()
            run(12, 4)
            print("rep %d world %d blocks/SM %d refill %2d: %.3f ms per frame" % (rep, world, bps, refill, (time.perf_counter() - t0) / 12 * 1e3), flush=True)
