"""Developer script: one b2rt_build_bvh call on the bench scene, for an ncu launch list of the build kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
lt, mats = prod.host.load_triangles(path)
with prod.Context(0) as ctx:
    ctx.build_bvh(lt[:4096])
    ctx.build_bvh(lt)
    print("build: %.3f s" % ctx.last_build_seconds)
