"""ctypes bindings to the CHECKERS (test infrastructure only):

  oracle/liboracle.so          our C restatement of kernel_bvh.cl          (oracle/rt_oracle.c)
  oracle/_ref/libref_oracle.so the reference's own sources compiled verbatim (oracle/build_ref.py)
  tests/emu/libemu.so          the product's traversal primitives + wide-BVH builder compiled for
                               the host, so their logic is testable without a GPU

Nothing here is reachable from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
CSRC = os.path.join(ROOT, "mini-opencl-raytracer_b200", "csrc")
EMU_DIR = os.path.join(ROOT, "tests", "emu")

RAY = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("tmin", "<f4"),
                ("dx", "<f4"), ("dy", "<f4"), ("dz", "<f4"), ("tmax", "<f4")])
HIT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<u4")])
REFHIT = np.dtype([("hit", "<i4"), ("tri", "<i4"), ("t", "<f4"), ("pos", "<f4", 3), ("normal", "<f4", 3),
                   ("uv", "<f4", 2)])
ATTR = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("uv", "<f4", 2)])
assert REFHIT.itemsize == 44


class OrCounters(C.Structure):
    _fields_ = [("nodes_visited", C.c_uint64), ("leaves_entered", C.c_uint64), ("tris_tested", C.c_uint64)]


class EmuStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_wide", "n_leaf_blocks", "leaf_words", "n_children", "max_depth_binary",
                                          "max_depth_wide", "stack_bound", "wide_visits", "leaf_blocks", "leaf_pass",
                                          "tri_tests", "words", "overflow", "max_stack")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_oracle():
    out = os.path.join(ORACLE_DIR, "liboracle.so")
    if _stale(out, [os.path.join(ORACLE_DIR, f) for f in ("rt_oracle.c", "rt_oracle.h", "Makefile")]):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "-B", "liboracle.so"])
    return out


def build_ref():
    """oracle/_ref: built from /root/reference where that exists, else the prebuilt .so that travelled."""
    out = os.path.join(ORACLE_DIR, "_ref", "libref_oracle.so")
    ref_dir = os.environ.get("B2RT_REFERENCE_DIR", "/root/reference")
    srcs = [os.path.join(ORACLE_DIR, "build_ref.py")] + \
           [os.path.join(ORACLE_DIR, "ref_build", f) for f in ("ref_host.cpp", "ref_kernel.cpp", "ref_api.h", "cl_shim.hpp")]
    if os.path.isdir(ref_dir) and _stale(out, srcs):
        subprocess.check_call(["python3", os.path.join(ORACLE_DIR, "build_ref.py"), "-q"])
    return out if os.path.exists(out) else None


def build_emu():
    out = os.path.join(EMU_DIR, "libemu.so")
    srcs = [os.path.join(EMU_DIR, "emu.cpp"), os.path.join(EMU_DIR, "warp_emu.cpp")] + \
           [os.path.join(CSRC, f) for f in ("traverse.cuh", "coop.cuh", "wide_bvh.cpp", "wide_bvh.h", "b2rt_types.h")]
    if _stale(out, srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-frounding-math",
                               "-fno-fast-math", "-pthread", "-DB2_EMU_CHECK_CULLING", "-x", "c++", "-I", CSRC, os.path.join(EMU_DIR, "emu.cpp"),
                               os.path.join(EMU_DIR, "warp_emu.cpp"), os.path.join(CSRC, "wide_bvh.cpp"), "-o", out])
    return out


def build_loader_hooks():
    """tests/emu/libloaderhooks.so: the product's OBJ loader source compiled with -DG3D_TEST_HOOKS, which exposes its number
    and face-token scanners one token at a time (the product library itself does not export them)."""
    out = os.path.join(EMU_DIR, "libloaderhooks.so")
    pkg = os.path.join(ROOT, "mini-opencl-raytracer_b200")
    src = os.path.join(pkg, "host", "obj_loader.cpp")
    if _stale(out, [src, os.path.join(pkg, "host", "glaze3d.h")]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-DG3D_TEST_HOOKS", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(pkg, "host"), src, "-L", pkg, "-lglaze3d", "-lb2rt", "-Wl,-rpath," + pkg, "-o", out])
    return out


def scan_triplet(token):
    """One `v/vt/vn` face token as the product's CLOBJloader reads it (sscanf "%d/%d/%d" semantics)."""
    if "hooks" not in _cache:
        _cache["hooks"] = C.CDLL(build_loader_hooks())
    out = (C.c_uint * 3)(0, 0, 0)
    _cache["hooks"].g3d_scan_triplet(token.encode(), out)
    return tuple(int(x) for x in out)


def parse_numbers(text, max_count=1 << 20):
    """The numbers of `text` exactly as the product's CLOBJloader reads them (scanf("%f") semantics)."""
    if "hooks" not in _cache:
        _cache["hooks"] = C.CDLL(build_loader_hooks())
    L = _cache["hooks"]
    L.g3d_parse_numbers.restype = C.c_int
    L.g3d_parse_numbers.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    out = np.empty(max_count, dtype=np.float32)
    n = L.g3d_parse_numbers(text.encode(), out.ctypes.data, max_count)
    return out[:n].copy()


_cache = {}


def oracle():
    if "oracle" not in _cache:
        L = C.CDLL(build_oracle())
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        L.oracle_trace_closest.argtypes = [vp, vp, vp, u64, vp, vp, C.POINTER(OrCounters), C.c_int]
        L.oracle_trace_any.argtypes = [vp, vp, vp, u64, vp, C.c_int]
        L.oracle_render.argtypes = [vp, vp, vp, vp, u32, u32, u32, i32, i32, C.c_float, vp, vp, vp, u64, u64, C.c_int]
        L.oracle_camera_rays.argtypes = [u32, u32, u32, vp, vp, vp, u64, u64, vp]
        L.oracle_hardware_threads.restype = C.c_int
        for f in (L.oracle_trace_closest, L.oracle_trace_any, L.oracle_render, L.oracle_camera_rays):
            f.restype = None
        _cache["oracle"] = L
    return _cache["oracle"]


def ref():
    if "ref" not in _cache:
        p = build_ref()
        if p is None:
            _cache["ref"] = None
        else:
            L = C.CDLL(p)
            vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
            L.ref_scene_load.restype = vp
            L.ref_scene_load.argtypes = [C.c_char_p, C.c_uint]
            L.ref_scene_from_triangles.restype = vp
            L.ref_scene_from_triangles.argtypes = [vp, u64, vp, u64, C.c_uint]
            L.ref_scene_free.argtypes = [vp]
            for n in ("ref_scene_num_triangles", "ref_scene_num_nodes", "ref_scene_num_materials"):
                getattr(L, n).restype = u64
                getattr(L, n).argtypes = [vp]
            for n in ("ref_scene_triangles", "ref_scene_nodes", "ref_scene_materials"):
                getattr(L, n).restype = vp
                getattr(L, n).argtypes = [vp]
            L.ref_intersect.argtypes = [vp, vp, vp, u64, vp, C.c_int]
            L.ref_intersect.restype = None
            L.ref_render.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, i32, i32, C.c_float, vp, vp, vp, u64, u64, C.c_int]
            L.ref_render.restype = None
            _cache["ref"] = L
    return _cache["ref"]


def emu():
    if "emu" not in _cache:
        L = C.CDLL(build_emu())
        L.emu_build.restype = C.c_char_p
        L.emu_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(EmuStats)]
        L.emu_trace.restype = None
        L.emu_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(EmuStats), C.c_uint32]
        _cache["emu"] = L
    return _cache["emu"]


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _v3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32)[:3])


# ---- scenes built by the reference's own loader + builder -------------------------------------
def ref_load_scene(obj_path, max_prims=4):
    """(tris, nodes, mats) as uint8 arrays of 256/48/64-byte records, built by the REFERENCE's
    CLOBJloader::Load + CLBVHScene::CreateBVHTrees (CLEngineBase.cpp:172-179)."""
    L = ref()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_oracle.so is not available")
    if len(os.fsencode(obj_path)) > 70:
        # the reference builds the .mtl name in an 80-byte buffer (CLOBJloader.cpp:18-23, 133-138): hand it a short path
        import shutil
        import tempfile
        d = tempfile.mkdtemp(prefix="ref")
        short = os.path.join(d, "s.obj")
        shutil.copy(obj_path, short)
        shutil.copy(obj_path[:-4] + ".mtl", short[:-4] + ".mtl")
        try:
            return ref_load_scene(short, max_prims)
        finally:
            shutil.rmtree(d, ignore_errors=True)
    s = L.ref_scene_load(os.fsencode(obj_path), max_prims)
    if not s:
        raise RuntimeError("reference loader failed on %s" % obj_path)
    try:
        out = []
        for count, data, size in ((L.ref_scene_num_triangles, L.ref_scene_triangles, 256),
                                  (L.ref_scene_num_nodes, L.ref_scene_nodes, 48),
                                  (L.ref_scene_num_materials, L.ref_scene_materials, 64)):
            n = int(count(s))
            buf = (C.c_uint8 * (n * size)).from_address(data(s)) if n else (C.c_uint8 * 0)()
            out.append(np.frombuffer(buf, dtype=np.uint8).copy().reshape(n, size))
        return tuple(out)
    finally:
        L.ref_scene_free(s)


# ---- tracing -------------------------------------------------------------------------------
def oracle_closest(tris, nodes, rays, threads=0, want_counters=False, want_attr=False):
    rays = np.ascontiguousarray(rays)
    hits = np.empty(rays.shape[0], dtype=HIT)
    attr = np.empty(rays.shape[0], dtype=ATTR) if want_attr else None
    cnt = OrCounters()
    oracle().oracle_trace_closest(_p(tris), _p(nodes), _p(rays), rays.shape[0], _p(hits), _p(attr),
                                  C.byref(cnt) if want_counters else None, threads)
    res = [hits]
    if want_counters:
        res.append({"nodes_visited": cnt.nodes_visited, "leaves_entered": cnt.leaves_entered, "tris_tested": cnt.tris_tested})
    if want_attr:
        res.append(attr)
    return res[0] if len(res) == 1 else tuple(res)


def oracle_any(tris, nodes, rays, threads=0):
    rays = np.ascontiguousarray(rays)
    occ = np.empty(rays.shape[0], dtype=np.uint32)
    oracle().oracle_trace_any(_p(tris), _p(nodes), _p(rays), rays.shape[0], _p(occ), threads)
    return occ


def ref_closest(tris, nodes, rays, threads=0):
    """The verbatim reference Intersect() (always tmax = 100000)."""
    rays = np.ascontiguousarray(rays)
    out = np.empty(rays.shape[0], dtype=REFHIT)
    ref().ref_intersect(_p(tris), _p(nodes), _p(rays), rays.shape[0], _p(out), threads)
    return out


def oracle_render(tris, nodes, mats, result, width, height, frame_count, bounces, light_type=0, sky=1.0,
                  pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0), gid0=0, gid1=None, threads=0):
    gid1 = width * height if gid1 is None else gid1
    a, b, c = _v3(pos), _v3(front), _v3(up)
    oracle().oracle_render(_p(tris), _p(nodes), _p(mats), _p(result), width, height, frame_count, bounces, light_type,
                           sky, _p(a), _p(b), _p(c), gid0, gid1, threads)
    return result


def ref_render(tris, nodes, mats, result, width, height, frame_count, bounces, light_type=0, sky=1.0,
               pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0), gid0=0, gid1=None, threads=0):
    gid1 = width * height if gid1 is None else gid1
    a, b, c = _v3(pos), _v3(front), _v3(up)
    ref().ref_render(_p(tris), _p(nodes), _p(mats), _p(result), width, height, frame_count, 0, bounces, light_type,
                     sky, _p(a), _p(b), _p(c), gid0, gid1, threads)
    return result


def oracle_camera_rays(width, height, frame_count, pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0),
                       gid0=0, gid1=None):
    gid1 = width * height if gid1 is None else gid1
    rays = np.empty(gid1 - gid0, dtype=RAY)
    a, b, c = _v3(pos), _v3(front), _v3(up)
    oracle().oracle_camera_rays(width, height, frame_count, _p(a), _p(b), _p(c), gid0, gid1, _p(rays))
    return rays


# ---- host emulation of the product's traversal ------------------------------------------------
def emu_culling_violations():
    """Children let through by the exact-arithmetic wide-node test (test_wide_node_robust) but culled by the fast one, counted
    over every emulated walk of this process: the fast test must be conservative, so this stays 0."""
    L = emu()
    L.emu_culling_violations.restype = C.c_uint64
    return int(L.emu_culling_violations())


def emu_build(tris, nodes):
    st = EmuStats()
    err = emu().emu_build(_p(nodes), nodes.shape[0], _p(tris), tris.shape[0], C.byref(st))
    if err:
        raise RuntimeError(err.decode())
    return st


def emu_trace_coop(rays, any_hit=False, stats=None, handoff=0, wide_limit=192, fcap=256, resume=0):
    """The warp-cooperative tail mode (csrc/coop.cuh) on 32 emulated lanes (tests/emu/warp_emu.cpp). handoff = 0: the
    whole ray runs cooperatively from the root; > 0: a pseudo-random number (< handoff) of solo steps first, then the
    solo lane's state is handed over like trace_persistent does when its ray pool runs dry. resume > 0: the two-step
    tail -- another lane restores the suspended state (Lane::resume), walks on alone (< resume steps) and suspends again."""
    rays = np.ascontiguousarray(rays)
    st = stats if stats is not None else EmuStats()
    L = emu()
    L.emu_trace_coop.restype = None
    L.emu_trace_coop.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    if any_hit:
        occ = np.empty(rays.shape[0], dtype=np.uint32)
        L.emu_trace_coop(_p(rays), rays.shape[0], None, _p(occ), 1, C.addressof(st), handoff, wide_limit, fcap, resume)
        return occ
    hits = np.empty(rays.shape[0], dtype=HIT)
    L.emu_trace_coop(_p(rays), rays.shape[0], _p(hits), None, 0, C.addressof(st), handoff, wide_limit, fcap, resume)
    return hits


def emu_trace_hybrid(rays, any_hit=False, depth=2, schedule=0):
    """The walk with the kernels' stack placement (first `depth` entries in a strided 'shared' column, the rest local) and
    the production build's branch-free push."""
    rays = np.ascontiguousarray(rays)
    L = emu()
    L.emu_trace_hybrid.restype = None
    L.emu_trace_hybrid.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
    if any_hit:
        occ = np.empty(rays.shape[0], dtype=np.uint32)
        L.emu_trace_hybrid(_p(rays), rays.shape[0], None, _p(occ), 1, depth, schedule)
        return occ
    hits = np.empty(rays.shape[0], dtype=HIT)
    L.emu_trace_hybrid(_p(rays), rays.shape[0], _p(hits), None, 0, depth, schedule)
    return hits


def emu_trace(rays, any_hit=False, stats=None, schedule=0):
    """schedule = 0: leaves are consumed as soon as they are queued; != 0: node and leaf steps are
    interleaved pseudo-randomly per ray (what the warp-vote scheduling of the kernels can produce)."""
    rays = np.ascontiguousarray(rays)
    st = stats if stats is not None else EmuStats()
    if any_hit:
        occ = np.empty(rays.shape[0], dtype=np.uint32)
        emu().emu_trace(_p(rays), rays.shape[0], None, _p(occ), 1, C.byref(st), schedule)
        return occ
    hits = np.empty(rays.shape[0], dtype=HIT)
    emu().emu_trace(_p(rays), rays.shape[0], _p(hits), None, 0, C.byref(st), schedule)
    return hits
