"""Developer timing of the frame path (megakernel): kernel-only vs RenderFrame with read-back."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
import scenes
prod = load_product()

def run(name, tris, nodes, mats, W, H, bounces, cam, modes=(0,)):
    with prod.Context(0) as ctx:
        ctx.upload_scene(tris, nodes, mats)
        ctx.resize(W, H)
        for mode in modes:
            ctx.set_option(prod.capi.OPT_RENDER_MODE, mode)
            ctx.set_frame(1, bounces, **cam); ctx.execute(W * H); ctx.finish()
            t0 = time.perf_counter()
            for f in range(2, 10):
                ctx.set_frame(f, bounces, **cam); ctx.execute(W * H)
            ctx.finish()
            k = (time.perf_counter() - t0) / 8 * 1e3
            px = np.empty((W * H, 4), dtype=np.float32)
            t0 = time.perf_counter()
            for f in range(10, 14):
                ctx.set_frame(f, bounces, **cam); ctx.execute(W * H); ctx.read_pixels(px)
            r = (time.perf_counter() - t0) / 4 * 1e3
            print("%-28s mode %d %dx%d b=%d: kernel %.3f ms/frame (%.1f Mpaths/s), with read-back %.3f ms" % (name, mode, W, H, bounces, k, W * H / k / 1e3, r), flush=True)

modes = tuple(int(x) for x in os.environ.get("MODES", "0").split(","))
t, n, m = prod.host.load_scene(scenes.CORNELL, 4)
run("cornell", t, n, m, 1920, 1080, 4, scenes.CAMERA, modes)
run("cornell", t, n, m, 1920, 1080, 9, scenes.CAMERA, modes)
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
t, n, m = prod.host.load_scene(path, 4)
cam = dict(pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, -0.3), up=(0.0, 0.0, 1.0))
run("icosphere 1M faces", t, n, m, 3840, 2160, 4, cam, modes)
