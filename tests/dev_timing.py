"""Developer timing script (not a test, not the bench): traversal throughput of the wide and binary
kernels on a displaced icosphere built by the reference's builder. Usage: python tests/dev_timing.py [subdiv] [nrays]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402
from conftest import load_product  # noqa: E402

prod = load_product()
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 7
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
d = tempfile.mkdtemp()
t0 = time.time()
p, n, f = scenes.displaced_sphere(subdiv)
path = scenes.write_obj(os.path.join(d, "s.obj"), p, n, f)
tris, nodes, mats = ol.ref_load_scene(path, 4)
print("scene: %d faces -> %d tris, %d nodes (%.1fs)" % (f.shape[0], tris.shape[0], nodes.shape[0], time.time() - t0), flush=True)
ctx = prod.Context(0)
t0 = time.time()
ctx.upload_scene(tris, nodes, mats)
print("upload+wide build %.2fs" % (time.time() - t0), ctx.scene_info(), flush=True)
rays = scenes.shell_rays(nrays, 10.0, seed=1)
d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
d_hits = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")
d_occ = torch.empty((nrays,), dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st = stream.cuda_stream


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for bps in (0, 4, 8, 12, 16):
    ctx.set_option(prod.capi.OPT_BLOCKS_PER_SM, bps)
    ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
    print("wide closest  blocks/SM=%2d  %.3f ms  %.1f Mrays/s" % (bps, ms, nrays / ms / 1e3), flush=True)
ctx.set_option(prod.capi.OPT_BLOCKS_PER_SM, 0)
for rm in (1, 4, 8, 16, 24):
    ctx.set_option(prod.capi.OPT_REFILL_MIN, rm)
    ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
    print("wide closest  refill_min=%2d  %.3f ms  %.1f Mrays/s" % (rm, ms, nrays / ms / 1e3), flush=True)
ctx.set_option(prod.capi.OPT_REFILL_MIN, 8)
for lb in (8, 12, 16, 24, 32, 64):
    ctx.set_option(prod.capi.OPT_LEAF_BIAS, lb)
    ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
    print("wide closest  leaf_bias=%2d/16  %.3f ms  %.1f Mrays/s" % (lb, ms, nrays / ms / 1e3), flush=True)
ctx.set_option(prod.capi.OPT_LEAF_BIAS, 16)
ms = timeit(lambda: ctx.trace_any_device(d_rays.data_ptr(), nrays, d_occ.data_ptr(), st))
print("wide any      %.3f ms  %.1f Mrays/s" % (ms, nrays / ms / 1e3))
ctx.set_option(prod.capi.OPT_TRAVERSAL, 1)
ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st), reps=2)
print("binary closest %.3f ms  %.1f Mrays/s" % (ms, nrays / ms / 1e3))
ctx.set_option(prod.capi.OPT_TRAVERSAL, 0)
ctx.set_option(prod.capi.OPT_COUNTERS, 1)
ctx.reset_counters()
ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st)
c = ctx.counters()
print("per ray:", {k: v / c["rays"] for k, v in c.items() if k != "rays"})
h = d_hits.cpu().numpy().view(prod.HIT_DTYPE).reshape(-1)
sub = slice(0, 200000)
want = ol.oracle_closest(tris, nodes, rays[sub])
print("parity on 200k: ids", (h[sub]["tri"] == want["tri"]).mean(), "t", (h[sub]["t"] == want["t"]).mean(), "hit frac", (want["tri"] != 0xFFFFFFFF).mean())
