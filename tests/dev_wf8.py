"""Developer script: one rank's 1/8 share (strided bands) of a 4K icosphere frame, wavefront, for an ncu launch list."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
W, H = 3840, 2160
cam = dict(pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, -0.3), up=(0.0, 0.0, 1.0))
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
with prod.Context(0) as ctx:
    ctx.upload_scene(t, n, m)
    ctx.resize(W, H)
    ctx.set_option(prod.capi.OPT_WAVEFRONT_LANES, int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    plan = prod.sharding.BandPlan(W, H, 8, band_rows=8)
    for f in (1, 2, 3):
        ctx.set_frame(f, 4, **cam)
        plan.render(ctx, 0)
    ctx.finish()
