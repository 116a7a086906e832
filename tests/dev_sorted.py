"""Developer experiment (not a test): how much does ray coherence buy? Times trace_closest on the bench scene with the same
rays unsorted and sorted on the host by Morton keys of several resolutions (origin + direction). Upper bound for an on-device
ray reorder stage. Usage: python tests/dev_sorted.py [nrays] [frequency]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_product  # noqa: E402

prod = load_product()
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
freq = int(sys.argv[2]) if len(sys.argv) > 2 else 224
path = "/tmp/b2rt_scenes/ico_f%d.obj" % freq
os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
if not os.path.exists(path):
    prod.host.write_icosphere_obj(path, freq, radius=10.0, amplitude=0.08, seed=7)
tris, nodes, mats = prod.host.load_scene(path, 4)
ctx = prod.Context(0)
ctx.upload_scene(tris, nodes, mats)
rays = prod.workloads.shell_rays(nrays, 10.0, seed=1000)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st = stream.cuda_stream
d_hits = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def spread(v, bits):
    out = np.zeros(v.shape, dtype=np.uint64)
    for b in range(bits):
        out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(6 * b)
    return out


def morton6(rays, bits):
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], 1).astype(np.float64)
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], 1).astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    lo, hi = o.min(0), o.max(0)
    qo = np.clip(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.uint64), 0, (1 << bits) - 1)
    qd = np.clip(((d * 0.5 + 0.5) * (1 << bits)).astype(np.uint64), 0, (1 << bits) - 1)
    key = np.zeros(rays.shape[0], dtype=np.uint64)
    for a in range(3):
        key |= spread(qo[:, a], bits) << np.uint64(2 * a + 1)
        key |= spread(qd[:, a], bits) << np.uint64(2 * a)
    return key


def run(label, r):
    d_rays = torch.from_numpy(np.ascontiguousarray(r).view(np.float32).reshape(-1, 8)).cuda()
    ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), r.shape[0], d_hits.data_ptr(), st))
    ms_any = timeit(lambda: ctx.trace_any_device(d_rays.data_ptr(), r.shape[0], d_hits.data_ptr(), st))
    ctx.set_option(prod.capi.OPT_COUNTERS, 1)
    ctx.reset_counters()
    ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), r.shape[0], d_hits.data_ptr(), st)
    torch.cuda.synchronize()
    c = ctx.counters()
    ctx.set_option(prod.capi.OPT_COUNTERS, 0)
    print("%-34s closest %8.3f ms %8.1f Mrays/s | any %8.1f Mrays/s | node lanes %.1f leaf lanes %.1f nodes/ray %.1f" % (
        label, ms, r.shape[0] / ms / 1e3, r.shape[0] / ms_any / 1e3, c["node_phase_lanes"] / max(c["node_phases"], 1),
        c["leaf_phase_lanes"] / max(c["leaf_phases"], 1), c["wide_nodes"] / max(c["rays"], 1)), flush=True)
    del d_rays


run("unsorted", rays)
for bits in (2, 3, 4, 5, 7):
    k = morton6(rays, bits)
    run("sorted, %d bits/dim (%d-bit key)" % (bits, 6 * bits), rays[np.argsort(k, kind="stable")])
# camera rays (fully coherent) for scale
ctx.resize(3840, 2160)
ctx.set_frame(1, 1, pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, -0.3), up=(0.0, 0.0, 1.0))
d = torch.empty((3840 * 2160, 8), dtype=torch.float32, device="cuda")
ctx.camera_rays_device(0, 3840 * 2160, d.data_ptr())
ctx.finish()
run("4K camera rays", d.cpu().numpy().view(prod.RAY_DTYPE).reshape(-1))
