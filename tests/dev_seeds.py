"""Developer check: closest-hit time of the bench scene for several ray-stream seeds."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
tris, nodes, mats = prod.host.load_scene(path, 4)
ctx = prod.Context(0)
ctx.upload_scene(tris, nodes, mats)
n = 1 << 24
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for seed in [int(a) for a in sys.argv[1:]] or [1000, 1001, 1002, 1003]:
    rays = prod.workloads.shell_rays(n, 10.0, seed=seed)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    for _ in range(2):
        ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ctx.set_option(prod.capi.OPT_COUNTERS, 1); ctx.reset_counters(); ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    c = ctx.counters(); ctx.set_option(prod.capi.OPT_COUNTERS, 0)
    print("   max steps/ray", c["max_steps_per_ray"], "mean nodes", c["wide_nodes"] / c["rays"])
    h = d_hits.cpu().numpy().view(prod.HIT_DTYPE).reshape(-1)
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], 1)
    print("seed %d: %.3f ms, hit frac %.4f, mean t %.3f, |o| mean %.3f" % (seed, e0.elapsed_time(e1) / 5, (h["tri"] != 0xFFFFFFFF).mean(), h["t"][h["tri"] != 0xFFFFFFFF].mean(), np.linalg.norm(o, axis=1).mean()), flush=True)
