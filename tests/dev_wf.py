"""Developer script: a few frames of the wavefront path on the 1M-face icosphere (4K) and cornell (1080p), for
`ncu --metrics gpu__time_duration.sum` launch lists. Usage: python tests/dev_wf.py [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
import scenes
prod = load_product()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
for name, obj, W, H, cam in (("ico", path, 3840, 2160, dict(pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, -0.3), up=(0.0, 0.0, 1.0))),
                             ("cornell", scenes.CORNELL, 1920, 1080, scenes.CAMERA)):
    t, n, m = prod.host.load_scene(obj, 4)
    with prod.Context(0) as ctx:
        ctx.upload_scene(t, n, m)
        ctx.resize(W, H)
        ctx.set_option(prod.capi.OPT_RENDER_MODE, mode)
        for f in (1, 2, 3):
            ctx.set_frame(f, 4, **cam)
            ctx.execute(W * H)
        ctx.finish()
    print(name, "done", flush=True)
