"""Developer script (GPU): where are the expensive primary rays of the 4K frame? max / mean solo steps per ray for every 8-row band,
then per row and per 64-pixel span of the worst band. usage: dev_hot.py [faces=10000000]"""
import os, sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
faces = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
W, H = 3840, 2160
path = "/tmp/b2rt_scenes/scatter_%d.obj" % faces
if not os.path.exists(path + ".done"):
    subprocess.check_call([os.path.join(os.path.dirname(prod.lib_path()), "scenegen"), "scattered", path, str(faces), "50.0", "0.05", "0.5", "11"], stdout=subprocess.DEVNULL)
    open(path + ".done", "w").close()
cam = dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
with prod.Context(0) as ctx:
    ctx.upload_scene(t, n, m)
    ctx.set_arg(cap.ARG_WIDTH, np.uint32(W)); ctx.set_arg(cap.ARG_HEIGHT, np.uint32(H))
    ctx.set_frame(1, 1, **cam)
    d_rays = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    d_hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
    ctx.camera_rays_device(0, W * H, d_rays.data_ptr())
    ctx.finish()
    ctx.set_option(cap.OPT_COOP_MAX, 0)

    def stats(g0, g1):
        ctx.set_option(cap.OPT_COUNTERS, 1)
        ctx.reset_counters()
        ctx.trace_closest_device(d_rays.data_ptr() + g0 * 32, g1 - g0, d_hits.data_ptr() + g0 * 16)
        c = ctx.counters()
        ctx.set_option(cap.OPT_COUNTERS, 0)
        return c["max_steps_per_ray"], (c["wide_nodes"] + c["leaf_blocks"]) / max(c["rays"], 1)

    bands = [(b,) + stats(b * 8 * W, min((b + 1) * 8 * W, W * H)) for b in range(H // 8)]
    worst = sorted(bands, key=lambda x: -x[1])[:8]
    print("bands by max steps/ray (band, max, mean):", [(b, mx, round(mean, 1)) for b, mx, mean in worst])
    print("median band max:", int(np.median([b[1] for b in bands])), "mean of means: %.1f" % np.mean([b[2] for b in bands]))
    b = worst[0][0]
    rows = [(y,) + stats(y * W, (y + 1) * W) for y in range(b * 8, b * 8 + 8)]
    print("rows of band %d (row, max, mean):" % b, [(y, mx, round(mean, 1)) for y, mx, mean in rows])
    y = max(rows, key=lambda x: x[1])[0]
    spans = [(x,) + stats(y * W + x, y * W + x + 64) for x in range(0, W, 64)]
    top = sorted(spans, key=lambda s: -s[1])[:6]
    print("64-pixel spans of row %d (x, max, mean):" % y, [(x, mx, round(mean, 1)) for x, mx, mean in top])
    x0 = top[0][0]
    px = [(x,) + stats(y * W + x, y * W + x + 1) for x in range(x0, x0 + 64)]
    print("pixels of the worst span (x, steps):", [(x, mx) for x, mx, _ in sorted(px, key=lambda s: -s[1])[:8]])
    xs = max(px, key=lambda s: s[1])[0]
    r = d_rays[y * W + xs].cpu().numpy()
    h = d_hits[y * W + xs].cpu().numpy()
    print("worst ray: origin %s dir %s -> t %.3f tri %d" % (r[:3], r[4:7], h[0], h[3:4].view(np.uint32)[0]))
    # the whole frame's primary rays as one stream: time with the tail mode off / on
    for coop in (0, 8):
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        for _ in range(2):
            ctx.trace_closest_device(d_rays.data_ptr(), W * H, d_hits.data_ptr())
        ctx.finish()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.trace_closest_device(d_rays.data_ptr(), W * H, d_hits.data_ptr())
        ctx.finish()
        print("primary rays, coop %d: %.3f ms" % (coop, (time.perf_counter() - t0) / 3 * 1e3))
