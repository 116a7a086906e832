"""Seeded synthetic scenes and ray sets for the parity tests (numpy only).

Scenes are written as OBJ + MTL text in the dialect CLOBJloader accepts (SURVEY.md
Appendix A-7: `v/vt/vn` triplets, sibling .mtl, one `usemtl`) and are then ingested by
the loader + BVH builder under test, so triangle IDs are defined the reference's way.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CORNELL = os.path.join(GOLDEN, "cornell.obj")   # the reference's bundled asset (data fixture, unedited)

CAMERA = dict(pos=(0.0, -25.0, 8.5), front=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0))   # CLcamera.h:8-10


def icosphere(subdiv):
    """Unit icosphere by recursive subdivision: (verts (n,3) f64, faces (m,3) i64), outward CCW."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        mid = {}

        def midpoint(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in mid:
                p = verts[key[0]] + verts[key[1]]
                verts.append(p / np.linalg.norm(p))
                mid[key] = len(verts) - 1
            return mid[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    return np.array(verts), np.array(faces, dtype=np.int64)


def displaced_sphere(subdiv, radius=10.0, amplitude=0.08, seed=7):
    """Bumpy closed mesh: icosphere displaced radially by a smooth seeded function. Returns
    (positions f32, normals f32, faces)."""
    rng = np.random.default_rng(seed)
    d, faces = icosphere(subdiv)
    k = rng.normal(size=(6, 3)) * 4.0
    ph = rng.uniform(0, 2 * np.pi, size=6)
    r = 1.0 + amplitude * np.sin(d @ k.T + ph).sum(axis=1) / 3.0
    p = d * (radius * r)[:, None]
    fn = np.cross(p[faces[:, 1]] - p[faces[:, 0]], p[faces[:, 2]] - p[faces[:, 0]])
    n = np.zeros_like(p)
    for c in range(3):
        np.add.at(n, faces[:, c], fn)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return p.astype(np.float32), n.astype(np.float32), faces


def scattered_triangles(n, extent=50.0, seed=11, emin=0.05, emax=0.5):
    """n small triangles with centres uniform in [-extent, extent]^3 and random orientation."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n, 3))
    e = rng.uniform(emin, emax, size=(n, 1, 1))
    offs = rng.normal(size=(n, 3, 3))
    offs /= np.linalg.norm(offs, axis=2, keepdims=True)
    p = (c[:, None, :] + offs * e).reshape(-1, 3)
    fn = np.cross(p[1::3] - p[0::3], p[2::3] - p[0::3])
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-30)
    normals = np.repeat(fn, 3, axis=0)
    faces = np.arange(3 * n, dtype=np.int64).reshape(n, 3)
    return p.astype(np.float32), normals.astype(np.float32), faces


def write_obj(path, positions, normals, faces, quads=None, kd=(0.7, 0.7, 0.7)):
    """OBJ + sibling MTL. `quads` (m,4) adds 4-gon faces (the loader fans them with its own rule)."""
    assert path.endswith(".obj") and len(os.path.basename(path)) <= 75
    base = os.path.basename(path)[:-4]
    with open(path[:-4] + ".mtl", "w") as f:
        f.write("newmtl surface\nNs 9999.0\nKd %g %g %g\nKs 0 0 0\nKe 0 0 0\nNi 1.0\n" % kd)
        f.write("newmtl lamp\nNs 9999.0\nKd 0.8 0.8 0.8\nKs 0 0 0\nKe 1 1 1\nNi 1.0\n")
    with open(path, "w") as f:
        f.write("mtllib %s.mtl\n" % base)
        f.write("".join("v %.9g %.9g %.9g\n" % tuple(p) for p in positions))
        f.write("vt 0 0\nvt 1 0\nvt 0 1\n")
        f.write("".join("vn %.9g %.9g %.9g\n" % tuple(n) for n in normals))
        f.write("usemtl surface\n")
        fi = np.asarray(faces) + 1
        f.write("".join("f %d/1/%d %d/2/%d %d/3/%d\n" % (a, a, b, b, c, c) for a, b, c in fi))
        if quads is not None:
            f.write("usemtl lamp\n")
            for q in np.asarray(quads) + 1:
                f.write("f " + " ".join("%d/1/%d" % (i, i) for i in q) + "\n")
    return path


# ---- ray sets ----------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def shell_rays(n, radius, seed=1, centre=(0.0, 0.0, 0.0), tmax=100000.0):
    """origins uniform on a sphere of radius 3R, targets uniform in the ball of
    radius R: incoherent, almost every ray hits a closed mesh of radius ~R."""
    from importlib import import_module  # noqa: F401  (keeps this module free of product imports)
    rng = np.random.default_rng(seed)
    o = _unit(rng.normal(size=(n, 3))) * (3.0 * radius)
    tgt = _unit(rng.normal(size=(n, 3))) * (radius * rng.uniform(0, 1, size=(n, 1)) ** (1.0 / 3.0))
    d = tgt - o
    return pack_rays(o + np.asarray(centre), d, tmax)


def box_rays(n, lo, hi, seed=2, tmax=100000.0):
    """Origins uniform in a box, directions uniform on the sphere (interior + grazing cases)."""
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, size=(n, 3))
    d = _unit(rng.normal(size=(n, 3)))
    return pack_rays(o, d, tmax)


def axis_rays(lo, hi, n_per_axis=64, seed=3, tmax=100000.0):
    """Axis-parallel rays (zero direction components -> +-inf inverse, 0*inf = NaN in RayBounds)."""
    rng = np.random.default_rng(seed)
    out = []
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    for a in range(3):
        for s in (-1.0, 1.0):
            o = rng.uniform(lo, hi, size=(n_per_axis, 3))
            o[:, a] = hi[a] + 5.0 if s < 0 else lo[a] - 5.0
            d = np.zeros((n_per_axis, 3))
            d[:, a] = s
            out.append(pack_rays(o, d, tmax))
    return np.concatenate(out)


def bounce_rays(rays, hits, normals_of_hit, seed=4, tmax=100000.0, jitter=0.6):
    """One cosine-weighted diffuse bounce per hit, origin = pos + wi*0.01 (kernel_bvh.cl:380):
    the incoherent secondary set. `jitter` perturbs the shading normal the way a rough
    normal-interpolated mesh does, so a share of the rays dips below the geometric facet and
    re-hits it at t ~ -0.01: the reference's negative-t acceptance."""
    rng = np.random.default_rng(seed)
    m = hits["tri"] != 0xFFFFFFFF
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], axis=1)[m].astype(np.float64)
    d = _unit(np.stack([rays["dx"], rays["dy"], rays["dz"]], axis=1)[m].astype(np.float64))
    pos = o + d * hits["t"][m, None]
    n = _unit(normals_of_hit[m].astype(np.float64))
    n = _unit(n + jitter * rng.normal(size=n.shape))
    a = np.where(np.abs(n[:, :1]) > 0.5, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t = _unit(np.cross(a, n))
    s = np.cross(n, t)
    phi = rng.uniform(0, 2 * np.pi, size=(n.shape[0], 1))
    r2 = rng.uniform(0, 1, size=(n.shape[0], 1))
    wi = _unit(s * np.cos(phi) * np.sqrt(r2) + t * np.sin(phi) * np.sqrt(r2) + n * np.sqrt(1 - r2))
    return pack_rays(pos + wi * 0.01, wi, tmax)


def pack_rays(o, d, tmax=100000.0):
    dt = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("tmin", "<f4"),
                   ("dx", "<f4"), ("dy", "<f4"), ("dz", "<f4"), ("tmax", "<f4")])
    r = np.zeros(o.shape[0], dtype=dt)
    o = o.astype(np.float32)
    d = d.astype(np.float32)
    r["ox"], r["oy"], r["oz"] = o[:, 0], o[:, 1], o[:, 2]
    r["dx"], r["dy"], r["dz"] = d[:, 0], d[:, 1], d[:, 2]
    r["tmax"] = tmax
    return r


def tri_normals(tris_u8, hits):
    """Geometric-ish shading normal of each hit from the 256-byte triangle records (for bounce_rays)."""
    f = tris_u8.view(np.float32).reshape(-1, 64)
    idx = np.where(hits["tri"] == 0xFFFFFFFF, 0, hits["tri"]).astype(np.int64)
    n1, n2, n3 = f[idx, 8:11], f[idx, 28:31], f[idx, 48:51]
    u, v = hits["u"][:, None], hits["v"][:, None]
    return n2 * u + n3 * v + n1 * (1 - u - v)


def psnr(a, b):
    """PSNR in dB of two float images compared on values clamped to [0,1]."""
    a = np.clip(np.nan_to_num(a.astype(np.float64), nan=0.0, posinf=1.0, neginf=0.0), 0, 1)
    b = np.clip(np.nan_to_num(b.astype(np.float64), nan=0.0, posinf=1.0, neginf=0.0), 0, 1)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(1.0 / mse)


def tie_grid(n=48, layers=3):
    """A scene made to produce exact `t` ties between DIFFERENT triangles in DIFFERENT leaves: `layers` coincident copies
    of an n x n lattice of unit squares in the plane z = 0 (each square two triangles, normals +z), plus a second set of
    copies at z = -1. Rays aimed at lattice points, edge midpoints and cell centres hit vertices / edges exactly, so up
    to 6 x layers x 2 candidates share one bit-identical t; the reference keeps the first one ITS walk reaches
    (strict `<`, kernel_bvh.cl:140), which only an order-faithful traversal reproduces. Also includes degenerate
    (zero-area) triangles, which every walk must reject (det < 1e-8, kernel_bvh.cl:116)."""
    xs = np.arange(n + 1, dtype=np.float32)
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    pos, faces = [], []
    for z in (0.0, -1.0):
        for _ in range(layers):
            base = len(pos) * (n + 1) * (n + 1)
            pos.append(np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, z, dtype=np.float32)], 1))
            idx = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).ravel() + base
            # counter-clockwise seen from +z: front faces for rays coming down the z axis
            faces.append(np.stack([idx, idx + (n + 1), idx + (n + 2)], 1))
            faces.append(np.stack([idx, idx + (n + 2), idx + 1], 1))
    pos = np.concatenate(pos).astype(np.float32)
    faces = np.concatenate(faces)
    degenerate = np.stack([np.arange(0, 200), np.arange(0, 200), np.arange(1, 201)], 1)      # two equal vertices
    faces = np.concatenate([faces, degenerate])
    normals = np.tile(np.array([[0.0, 0.0, 1.0]], dtype=np.float32), (pos.shape[0], 1))
    return pos, normals, faces


def tie_rays(n=48, seed=8):
    """Straight-down and slanted rays through lattice points, edge midpoints and cell centres of tie_grid(n)."""
    rng = np.random.default_rng(seed)
    half = np.arange(0, 2 * n + 1, dtype=np.float64) * 0.5
    tx, ty = np.meshgrid(half, half, indexing="ij")
    tgt = np.stack([tx.ravel(), ty.ravel(), np.zeros(tx.size)], 1)
    down = pack_rays(tgt + np.array([0.0, 0.0, 8.0]), np.tile(np.array([[0.0, 0.0, -1.0]]), (tgt.shape[0], 1)))
    # slanted: origins with coordinates that are exact in binary, so that many of these still hit vertices exactly
    o = np.stack([rng.integers(0, n + 1, tgt.shape[0]), rng.integers(0, n + 1, tgt.shape[0]), np.full(tgt.shape[0], 16)], 1).astype(np.float64)
    slant = pack_rays(o, tgt - o)
    return np.concatenate([down, slant])


def moller_trumbore_all(tris_u8, rays):
    """RayTriangle (kernel_bvh.cl:98-153) for every (ray, triangle) pair in numpy float32, one rounding per operation
    in the reference's order: returns (accept, t) of shape (rays, tris) with accept = every test except `t < best`."""
    f = tris_u8.view(np.float32).reshape(-1, 64)
    v1, v2, v3 = f[None, :, 0:3], f[None, :, 20:23], f[None, :, 40:43]
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], 1).astype(np.float32)[:, None, :]
    d = np.stack([rays["dx"], rays["dy"], rays["dz"]], 1).astype(np.float32)
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = (d / ln[:, None]).astype(np.float32)[:, None, :]

    def dot(a, b):
        return ((a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]).astype(np.float32)

    def cross(a, b):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(np.float32)
    with np.errstate(all="ignore"):
        e1, e2 = v2 - v1, v3 - v1
        pvec = cross(np.broadcast_to(d, (d.shape[0], f.shape[0], 3)), np.broadcast_to(e2, (d.shape[0], f.shape[0], 3)))
        det = dot(np.broadcast_to(e1, pvec.shape), pvec)
        inv = (np.float32(1.0) / det).astype(np.float32)
        tvec = (o - v1).astype(np.float32)
        u = (dot(tvec, pvec) * inv).astype(np.float32)
        qvec = cross(tvec, np.broadcast_to(e1, tvec.shape))
        v = (dot(np.broadcast_to(d, qvec.shape), qvec) * inv).astype(np.float32)
        t = (dot(np.broadcast_to(e2, qvec.shape), qvec) * inv).astype(np.float32)
        ok = (det >= np.float32(1e-8)) & (u >= 0) & (u <= 1) & (v >= 0) & ((u + v).astype(np.float32) <= 1)
    return ok, t


def fuzz_scene(case, rng):
    """Small random scenes of awkward shapes, by `case % 6`: uniform soup, slivers, two distant clusters, a few huge
    triangles over many tiny ones, an axis-aligned sheet on an exact lattice, many exact duplicates. Every 17th face has
    zero area. Returns (positions, faces)."""
    k = int(rng.integers(20, 400))
    kind = case % 6
    if kind == 0:                                             # uniform soup
        c = rng.uniform(-5, 5, (k, 1, 3)); e = rng.normal(0, 1.0, (k, 3, 3))
    elif kind == 1:                                           # slivers: one long edge, tiny height
        c = rng.uniform(-5, 5, (k, 1, 3)); e = rng.normal(0, 1.0, (k, 3, 3)) * np.array([3.0, 0.01, 0.01])
    elif kind == 2:                                           # two tight clusters far apart
        c = np.where(rng.random((k, 1, 1)) < 0.5, -40.0, 40.0) + rng.normal(0, 0.2, (k, 1, 3)); e = rng.normal(0, 0.1, (k, 3, 3))
    elif kind == 3:                                           # a few huge triangles over many tiny ones
        c = rng.uniform(-5, 5, (k, 1, 3)); e = rng.normal(0, 0.05, (k, 3, 3)); e[:5] *= 400.0
    elif kind == 4:                                           # axis-aligned sheet at z = 0 on an exact lattice
        c = np.concatenate([rng.integers(-8, 8, (k, 1, 2)).astype(np.float64), np.zeros((k, 1, 1))], 2)
        e = np.concatenate([rng.integers(-2, 3, (k, 3, 2)).astype(np.float64), np.zeros((k, 3, 1))], 2)
    else:                                                     # many exact duplicates of a handful of triangles
        base_c = rng.uniform(-3, 3, (6, 1, 3)); base_e = rng.normal(0, 1.0, (6, 3, 3))
        pick = rng.integers(0, 6, k); c, e = base_c[pick], base_e[pick]
    pos = (c + e).reshape(-1, 3).astype(np.float32)
    faces = np.arange(3 * k).reshape(k, 3)
    faces[:: 17, 1] = faces[:: 17, 0]                          # a few zero-area faces (two equal vertices)
    return pos, faces
