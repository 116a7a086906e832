"""Developer script (GPU): stream rate on the bench scene under scheduling options (leaf vote weight, refill threshold).
usage: dev_sweep.py [rays=50000000]"""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product
prod = load_product()
cap = prod.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
path = "/tmp/b2rt_scenes/ico_f224.obj"
if not os.path.exists(path + ".done"):
    subprocess.check_call([os.path.join(os.path.dirname(prod.host.lib_path()), "scenegen"), "icosphere", path, "224", "10.0", "0.08", "7"], stdout=subprocess.DEVNULL)
    open(path + ".done", "w").close()
t, nn, m = prod.host.load_scene(path, 4, cache=True)[:3]
rays = prod.workloads.shell_rays(n, 10.0, seed=1000)
d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
with prod.Context(0) as ctx:
    ctx.upload_scene(t, nn, m)

    def rate(reps=4):
        for _ in range(2):
            ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
        ctx.finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            e0.record()
            for _ in range(reps):
                ctx.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), st.cuda_stream)
            e1.record()
        torch.cuda.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6

    rate()
    for bias in (16, 24, 32, 40, 48, 64):
        ctx.set_option(cap.OPT_LEAF_BIAS, bias)
        print("leaf_bias %3d: %.0f Mrays/s" % (bias, rate()), flush=True)
    ctx.set_option(cap.OPT_LEAF_BIAS, 32)
    for rm in (4, 6, 8, 12, 16):
        ctx.set_option(cap.OPT_REFILL_MIN, rm)
        print("refill_min %2d: %.0f Mrays/s" % (rm, rate()), flush=True)
    ctx.set_option(cap.OPT_REFILL_MIN, 8)
    for bps in (6, 7, 8):
        ctx.set_option(cap.OPT_BLOCKS_PER_SM, bps)
        print("blocks/SM %d: %.0f Mrays/s" % (bps, rate()), flush=True)
