"""Developer measurement (not a test): device BVH build vs the host SAH build on the bench scene -- build time and
closest-hit throughput of the resulting trees on the same ray stream."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
arg = sys.argv[1] if len(sys.argv) > 1 else "224"
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if arg.startswith("scatter:"):
    path = "/tmp/b2rt_scenes/scatter_%s.obj" % arg[8:]
    if not os.path.exists(path):
        prod.host.write_scattered_obj(path, int(arg[8:]), extent=50.0, edge_min=0.05, edge_max=0.5, seed=11)
    RADIUS = 50.0
else:
    path = "/tmp/b2rt_scenes/ico_f%s.obj" % arg
    if not os.path.exists(path):
        prod.host.write_icosphere_obj(path, int(arg), radius=10.0, amplitude=0.08, seed=7)
    RADIUS = 10.0
t0 = time.perf_counter(); loader_tris, mats = prod.host.load_triangles(path); t_parse = time.perf_counter() - t0
t0 = time.perf_counter(); sah = prod.host.build_scene(loader_tris, mats, 4); t_sah = time.perf_counter() - t0
n = 1 << 23
rays = prod.workloads.shell_rays(n, RADIUS, seed=1000)
d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()

def rate(ctx):
    with torch.cuda.stream(st):
        for _ in range(2):
            ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), st.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    return n * 4 / e0.elapsed_time(e1) / 1e3

with prod.Context(0) as ctx:
    ctx.build_bvh(loader_tris[:1000])                                   # warm-up: module load, allocator
    t0 = time.perf_counter(); tris, nodes, order = ctx.build_bvh(loader_tris); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    nn = np.zeros((2 * loader_tris.shape[0] - 1, 48), dtype=np.uint8); oo = np.empty(loader_tris.shape[0], dtype=np.uint32)
    import ctypes as C
    k = C.c_uint64(0)
    ctx._ck(ctx._L.b2rt_build_bvh(ctx._h, loader_tris.ctypes.data, loader_tris.shape[0], nn.ctypes.data, nn.shape[0], C.byref(k), oo.ctypes.data))
    t_abi = time.perf_counter() - t0
    print("parse %.2f s | host SAH build (all cores, incl. copies) %.2f s | b2rt_build_bvh %.3f s (+ triangle re-order in numpy: %.3f s total)" % (t_parse, t_sah, t_abi, t_dev))
    ctx.upload_scene(tris, nodes, mats)
    info = ctx.scene_info()
    r_dev = rate(ctx)
    ctx.set_option(prod.capi.OPT_COUNTERS, 1); ctx.reset_counters(); ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), 0); torch.cuda.synchronize()
    c = ctx.counters(); ctx.set_option(prod.capi.OPT_COUNTERS, 0)
    print("device-built tree: %d nodes, %d wide nodes, %.1f Mrays/s, %.2f wide nodes + %.2f leaves per ray" % (
        nodes.shape[0], info["n_wide_nodes"], r_dev, c["wide_nodes"] / c["rays"], c["leaf_blocks"] / c["rays"]))
    h_dev = d_hits.cpu().numpy().view(prod.HIT_DTYPE).reshape(-1).copy()
with prod.Context(0) as ctx:
    ctx.upload_scene(*sah)
    info = ctx.scene_info()
    r_sah = rate(ctx)
    ctx.set_option(prod.capi.OPT_COUNTERS, 1); ctx.reset_counters(); ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), 0); torch.cuda.synchronize()
    c = ctx.counters(); ctx.set_option(prod.capi.OPT_COUNTERS, 0)
    print("SAH tree (reference builder): %d nodes, %d wide nodes, %.1f Mrays/s, %.2f wide nodes + %.2f leaves per ray" % (
        sah[1].shape[0], info["n_wide_nodes"], r_sah, c["wide_nodes"] / c["rays"], c["leaf_blocks"] / c["rays"]))
    h_sah = d_hits.cpu().numpy().view(prod.HIT_DTYPE).reshape(-1)
hit = h_sah["tri"] != 0xFFFFFFFF
print("same hit/miss: %.6f, t within 1e-4 rel: %.6f" % ((hit == (h_dev["tri"] != 0xFFFFFFFF)).mean(),
      (np.abs(h_dev["t"][hit].astype(np.float64) - h_sah["t"][hit]) <= 1e-4 * np.abs(h_sah["t"][hit])).mean()))
