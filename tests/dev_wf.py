"""Developer script: a few frames of the frame path on the 1M-face icosphere (4K), a scattered-triangle scene (4K) and
cornell (1080p), for `ncu --metrics gpu__time_duration.sum` launch lists. Usage: python tests/dev_wf.py [mode] [scenes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
import scenes
prod = load_product()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["ico", "scatter", "cornell"]
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
path = "/tmp/b2rt_scenes/ico_f224.obj"
if "ico" in which and not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
spath = "/tmp/b2rt_scenes/scatter_2000000.obj"
if "scatter" in which and not os.path.exists(spath):
    prod.host.write_scattered_obj(spath, 2000000, extent=50.0, edge_min=0.05, edge_max=0.5, seed=11)
for name, obj, W, H, cam in (("ico", path, 3840, 2160, dict(pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, -0.3), up=(0.0, 0.0, 1.0))),
                             ("scatter", spath, 3840, 2160, dict(pos=(0.0, -140.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))),
                             ("cornell", scenes.CORNELL, 1920, 1080, scenes.CAMERA)):
    if name not in which:
        continue
    t, n, m = prod.host.load_scene(obj, 4, cache=True)[:3]
    with prod.Context(0) as ctx:
        ctx.upload_scene(t, n, m)
        ctx.resize(W, H)
        ctx.set_option(prod.capi.OPT_RENDER_MODE, mode)
        for lanes in ((1, 2) if mode == 0 else (1,)):
            ctx.set_option(prod.capi.OPT_WAVEFRONT_LANES, lanes)
            for frac in (1, 2, 8):                                # the whole frame, and one rank's share of a 2- / 8-GPU frame
                plan = prod.sharding.BandPlan(W, H, frac, band_rows=8)
                for f in range(5 if mode == 2 else 1):             # mode 2 settles on a choice after four launches of a shape
                    ctx.set_frame(1, 4, **cam)
                    plan.render(ctx, 0)
                ctx.finish()
                t0 = time.perf_counter()
                for f in (2, 3, 4, 5):
                    ctx.set_frame(f, 4, **cam)
                    plan.render(ctx, 0)
                ctx.finish()
                print(name, "mode", mode, "lanes", lanes, "pixels 1/%d" % frac, "%.3f ms/frame" % ((time.perf_counter() - t0) / 4 * 1e3), flush=True)
