"""Developer script (GPU): a few frames of one rank's share of a tiled 4K frame, for ncu launch lists / captures.
usage: dev_share.py [faces=10000000] [coop=8] [lanes=1] [frames=3] [world=8] [mode=0]"""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
a = [int(x) for x in sys.argv[1:]] + [None] * 6
faces, coop, lanes, frames, world, mode = (a[0] if a[0] is not None else 10000000, a[1] if a[1] is not None else 8, a[2] or 1, a[3] or 3, a[4] or 8, a[5] or 0)
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
W, H = 3840, 2160
if faces > 0:
    path = "/tmp/b2rt_scenes/scatter_%d.obj" % faces
    if not os.path.exists(path):
        subprocess.check_call([os.path.join(os.path.dirname(prod.lib_path()), "scenegen"), "scattered", path, str(faces), "50.0", "0.05", "0.5", "11"], stdout=subprocess.DEVNULL)
    cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
else:
    import scenes
    path, cam = scenes.CORNELL, {}
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
with prod.Context(0) as ctx:
    ctx.upload_scene(t, n, m)
    ctx.resize(W, H)
    ctx.set_option(cap.OPT_RENDER_MODE, mode)
    ctx.set_option(cap.OPT_COOP_MAX, coop)
    ctx.set_option(cap.OPT_WAVEFRONT_LANES, lanes)
    plan = prod.sharding.BandPlan(W, H, world, band_rows=8)
    t0 = time.perf_counter()
    for f in range(1, frames + 1):
        ctx.set_frame(f, 4, **cam)
        if world > 1:
            plan.render(ctx, 0)
        else:
            ctx.execute(W * H)
    ctx.finish()
    print("%d frames: %.3f ms each (incl. first-use allocations)" % (frames, (time.perf_counter() - t0) / frames * 1e3))
