"""Developer script (GPU): per-stage device times and counters of chosen ranks' shares of the tiled 4K frame.
usage: dev_tail3.py [faces=10000000] [world=8] [ranks=0,3,7]"""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
faces = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ranks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,3,7").split(",")]
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
    ctx.set_option(cap.OPT_WAVEFRONT_LANES, 1)
    plan = prod.sharding.BandPlan(W, H, world, band_rows=8)
    for coop, resume, help_ in [tuple(int(y) for y in (x.split(":") + ["1"])[:3]) for x in os.environ.get("COOPS", "8:0:1,8:0:0,8:16,4:16,4:12,2:16").split(",")]:
        ctx.set_option(cap.OPT_COOP_MAX, coop)
        ctx.set_option(cap.OPT_RESUME_MAX, resume)
        ctx.set_option(cap.OPT_TAIL_HELP, help_)
        print("== coop %d resume %d help %d" % (coop, resume, help_))
        for r in ranks:
            ctx.set_option(cap.OPT_STAGE_TIMES, 1)
            for f in (1, 2, 3):
                ctx.set_frame(f, 4, **cam); plan.render(ctx, r); ctx.finish()
            st = ctx.stage_times()
            ctx.set_option(cap.OPT_STAGE_TIMES, 0)
            ctx.set_option(cap.OPT_COUNTERS, 1)
            ctx.reset_counters()
            ctx.set_frame(3, 4, **cam); plan.render(ctx, r); ctx.finish()
            c = ctx.counters()
            ctx.set_option(cap.OPT_COUNTERS, 0)
            print("coop %d resume %d rank %d stages (ms): %s | sum %.3f" % (coop, resume, r, " ".join("%s %.3f" % (k[:2], v) for k, v in st), sum(v for _, v in st)))
            print("      ", {k: c[k] for k in c if k in ("rays", "wide_nodes", "leaf_blocks", "max_steps_per_ray", "coop_rays", "coop_steps", "coop_max_rounds", "coop_max_steps", "resumed_rays")}, flush=True)
