"""Developer timing of the host-buffer ray-stream entry (b2rt_trace_closest) against plain PCIe copies."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
path = "/tmp/b2rt_scenes/ico_f224.obj"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, 224, radius=10.0, amplitude=0.08, seed=7)
t, nn, m = prod.host.load_scene(path, 4, cache=True)[:3]
n = 50_000_000
host_rays = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
rays = host_rays.numpy().view(prod.RAY_DTYPE).reshape(-1)
prod.workloads.shell_rays(n, 10.0, seed=1000, out=rays)
host_hits = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
hits = host_hits.numpy().view(prod.HIT_DTYPE).reshape(-1)
d = torch.empty((n, 8), dtype=torch.float32, device="cuda")
dh = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for _ in range(2):
    d.copy_(host_rays, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(host_rays, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D alone: %.1f GB/s" % (n * 32 / dt / 1e9))
t0 = time.perf_counter(); host_hits.copy_(dh, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H alone: %.1f GB/s" % (n * 16 / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(host_rays, non_blocking=True)
with torch.cuda.stream(s2): host_hits.copy_(dh, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D + D2H together: %.1f ms -> %.0f Mrays/s bound" % (dt * 1e3, n / dt / 1e6))
for chunk in (1 << 20, 1 << 22, 1 << 24):
    os.environ["B2RT_STREAM_CHUNK"] = str(chunk)
    with prod.Context(0) as ctx:
        ctx.upload_scene(t, nn, m)
        ctx.trace_closest(rays, hits)
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.trace_closest(rays, hits)
        dt = (time.perf_counter() - t0) / 3
        print("chunk %d: %.1f ms, %.0f Mrays/s" % (chunk, dt * 1e3, n / dt / 1e6), flush=True)
