"""Developer script (GPU): ray-stream throughput on the bench scene under option settings (A/B in one process).
usage: dev_stream.py [rays=33554432]"""
import os, sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
path = "/tmp/b2rt_scenes/ico_f224.obj"
if not os.path.exists(path + ".done"):
    subprocess.check_call([os.path.join(os.path.dirname(prod.lib_path()), "scenegen"), "icosphere", path, "224", "10.0", "0.08", "7"], stdout=subprocess.DEVNULL)
    open(path + ".done", "w").close()
t, nn, m = prod.host.load_scene(path, 4, cache=True)[:3]
rays = prod.workloads.shell_rays(n, 10.0, seed=1000)
d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
ref = None
with prod.Context(0) as ctx:
    ctx.upload_scene(t, nn, m)

    def rate(any_hit=False, reps=5):
        fn = ctx.trace_any_device if any_hit else ctx.trace_closest_device
        for _ in range(2):
            fn(d_rays.data_ptr(), n, d_hits.data_ptr())
        ctx.finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            e0.record()
            for _ in range(reps):
                fn(d_rays.data_ptr(), n, d_hits.data_ptr(), st.cuda_stream)
            e1.record()
        torch.cuda.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6

    for l2 in (1,):
        for coop in (8, 0, 8, 0, 4):
            try:
                ctx.set_option(cap.OPT_L2_PERSIST, l2)
                ctx.set_option(cap.OPT_COOP_MAX, coop)
            except prod.B2RTError:
                pass                          # an older library under B2RT_LIB
            r = rate()
            h = d_hits.clone()
            same = True if ref is None else bool(torch.equal(ref.view(torch.int32), h.view(torch.int32)))
            ref = h
            print("l2_persist %d coop %d: closest %.0f Mrays/s, any-hit %.0f Mrays/s, identical hits: %s" % (l2, coop, r, rate(True), same), flush=True)
