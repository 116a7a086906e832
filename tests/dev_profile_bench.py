"""Developer profiling target (not a test): launches of the closest-hit stream kernel on the BENCH workload
(1 003 520-face icosphere, shell rays). Usage: python tests/dev_profile_bench.py [nrays] [launches]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
t, n, m = prod.host.load_scene(path, 4, cache=True)[:3]
ctx = prod.Context(0)
ctx.upload_scene(t, n, m)
host = torch.empty((nrays, 8), dtype=torch.float32, pin_memory=True)
prod.workloads.shell_rays(nrays, 10.0, seed=1000, out=host.numpy().view(prod.RAY_DTYPE).reshape(-1))
d_rays = host.cuda()
d_out = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    for i in range(launches):
        if i == launches - 1:
            e0.record()
        ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_out.data_ptr(), stream.cuda_stream)
    e1.record()
torch.cuda.synchronize()
print("last launch %.3f ms, %.1f Mrays/s" % (e0.elapsed_time(e1), nrays / e0.elapsed_time(e1) / 1e3))
