"""Developer A/B harness (not a test): time several builds of libb2rt (variants/*.so) on the same scene and
rays inside one process-per-variant loop. Usage: python tests/dev_ab.py subdiv nrays name1 name2 ... [--opt k=v ...]"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import tempfile
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as ol
    import scenes
    from conftest import load_product
    prod = load_product()
    root_pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mini-opencl-raytracer_b200")
    subdiv, nrays = sys.argv[2], int(sys.argv[3])
    opts = [a.split("=") for a in sys.argv[4:]]
    d = tempfile.mkdtemp()
    cache = "/tmp/ab_scene_%s.npz" % subdiv
    cam = None
    if subdiv == "sph":                             # configs[2] of bench.py: 196 scattered spheres, 4K camera rays (coherent primary batch)
        os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
        path = "/tmp/b2rt_scenes/spheres196_f16.obj"
        if not os.path.exists(path + ".done"):
            subprocess.check_call([os.path.join(root_pkg, "scenegen"), "spheres", path, "196", "16", "30.0", "4.0", "0.05", "5"], stdout=subprocess.DEVNULL)
            open(path + ".done", "w").close()
        tris, nodes, mats, _ = prod.host.load_scene(path, 4, cache=True)
        cam = (3840, 2160)
        nrays = cam[0] * cam[1]
    elif subdiv.startswith("f"):                      # the bench scene: geodesic frequency, e.g. f224
        os.makedirs("/tmp/b2rt_scenes", exist_ok=True)
        path = "/tmp/b2rt_scenes/ico_%s.obj" % subdiv
        if not os.path.exists(path):
            prod.host.write_icosphere_obj(path, int(subdiv[1:]), radius=10.0, amplitude=0.08, seed=7)
        tris, nodes, mats, _ = prod.host.load_scene(path, 4, cache=True)
    elif os.path.exists(cache):
        z = np.load(cache)
        tris, nodes, mats = z["t"], z["n"], z["m"]
    else:
        p, n, f = scenes.displaced_sphere(int(subdiv))
        tris, nodes, mats = ol.ref_load_scene(scenes.write_obj(os.path.join(d, "s.obj"), p, n, f), 4)
        np.savez(cache, t=tris, n=nodes, m=mats)
    ctx = prod.Context(0)
    ctx.upload_scene(tris, nodes, mats)
    if cam:
        ctx.set_arg(prod.capi.ARG_WIDTH, np.uint32(cam[0]))
        ctx.set_arg(prod.capi.ARG_HEIGHT, np.uint32(cam[1]))
        ctx.set_frame(1, 1, pos=(0.0, -95.0, 0.0), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))
        d_rays = torch.empty((nrays, 8), dtype=torch.float32, device="cuda")
        ctx.camera_rays_device(0, nrays, d_rays.data_ptr(), 0)
        ctx.finish()
        torch.cuda.synchronize()
        rays = d_rays.cpu().numpy().view(prod.RAY_DTYPE).reshape(-1)
    else:
        rays = prod.workloads.shell_rays(nrays, 10.0, seed=1) if subdiv.startswith("f") else scenes.shell_rays(nrays, 10.0, seed=1)
        d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream

    def timeit(fn, reps=10 if nrays > (1 << 24) else 40):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    sweeps = [[]]
    for k, v in opts:
        sweeps = [s + [(k, x)] for s in sweeps for x in v.split(",")]
    for _ in range(12 if not cam else 60):            # clocks up before the first figure (the first measurement of a process read 5-10 % low)
        ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st)
    torch.cuda.synchronize()
    for sw in sweeps:
        for k, x in sw:
            ctx.set_option(getattr(prod.capi, "OPT_" + k.upper()), int(x))
        ms = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
        ms_any = timeit(lambda: ctx.trace_any_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
        ms2 = timeit(lambda: ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st))
        print("  %-28s closest %7.3f ms %7.1f / %7.1f Mrays/s | any %7.1f Mrays/s" % (" ".join("%s=%s" % kv for kv in sw), ms, nrays / ms / 1e3, nrays / ms2 / 1e3, nrays / ms_any / 1e3), flush=True)
    ctx.set_option(prod.capi.OPT_COUNTERS, 1)
    ctx.reset_counters()
    ctx.finish()
    ctx.trace_closest_device(d_rays.data_ptr(), nrays, d_hits.data_ptr(), st)
    torch.cuda.synchronize()
    c = ctx.counters()
    r = max(c["rays"], 1)
    print("  per ray: nodes %.2f leaves %.2f gate-pass %.2f tris %.2f bytes %.0f | node phases: %.1f lanes avg, leaf phases: %.1f lanes avg, node:leaf phases %.2f, refill %.1f lanes avg" % (
        c["wide_nodes"] / r, c["leaf_blocks"] / r, c["leaf_gate_pass"] / r, c["tri_tests"] / r, c["bytes_fetched"] / r,
        c["node_phase_lanes"] / max(c["node_phases"], 1), c["leaf_phase_lanes"] / max(c["leaf_phases"], 1),
        c["node_phases"] / max(c["leaf_phases"], 1), c["refill_lanes"] / max(c["refills"], 1)))
    h = d_hits.cpu().numpy().view(prod.HIT_DTYPE).reshape(-1)[:50000]
    want = ol.oracle_closest(tris, nodes, rays[:50000])
    print("  parity 50k: ids %.6f t %.6f" % ((h["tri"] == want["tri"]).mean(), (h["t"] == want["t"]).mean()))
    sys.exit(0)

subdiv, nrays = sys.argv[1], sys.argv[2]
names = [a for a in sys.argv[3:] if "=" not in a]
opts = [a for a in sys.argv[3:] if "=" in a]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(int(os.environ.get("DEV_AB_PASSES", "2"))):
    for name in names:
        lib = os.path.join(root, "mini-opencl-raytracer_b200", "libb2rt.so") if name == "default" else os.path.join(root, "variants", "libb2rt_%s.so" % name)
        print("[%s] pass %d" % (name, rep), flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", subdiv, nrays] + opts, env=dict(os.environ, B2RT_LIB=lib))
