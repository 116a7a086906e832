"""Developer profiling target (not a test): N launches of the closest-hit (and optionally any-hit) stream
kernel on a displaced icosphere. Usage: python tests/dev_profile.py [subdiv] [nrays] [launches] [any]"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402
from conftest import load_product  # noqa: E402

prod = load_product()
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 7
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
any_hit = len(sys.argv) > 4 and sys.argv[4] == "any"
d = tempfile.mkdtemp()
p, n, f = scenes.displaced_sphere(subdiv)
tris, nodes, mats = ol.ref_load_scene(scenes.write_obj(os.path.join(d, "s.obj"), p, n, f), 4)
ctx = prod.Context(0)
ctx.upload_scene(tris, nodes, mats)
rays = scenes.shell_rays(nrays, 10.0, seed=1)
d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
d_out = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    for i in range(launches):
        if i == launches - 1:
            e0.record()
        if any_hit:
            ctx.trace_any_device(d_rays.data_ptr(), nrays, d_out.data_ptr(), stream.cuda_stream)
        else:
            ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_out.data_ptr(), stream.cuda_stream)
    e1.record()
torch.cuda.synchronize()
if os.environ.get("B2_COUNT"):
    ctx.set_option(prod.capi.OPT_COUNTERS, 1)
    ctx.reset_counters()
    ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    c = ctx.counters()
    print({k: v / c["rays"] for k, v in c.items()})
print("last launch %.3f ms, %.1f Mrays/s" % (e0.elapsed_time(e1), nrays / e0.elapsed_time(e1) / 1e3))
