"""Seeded synthetic ray sets of the shapes BASELINE.json names (numpy, host side).
SURVEY.md 8d: (5) ray stream = origins uniform on a sphere of radius 3R, targets uniform in the ball
of radius R; (3) incoherent diffuse set = one cosine-weighted bounce per primary hit, origin
pos + wi*0.01 (kernel_bvh.cl:380)."""
import numpy as np

from .layouts import RAY_DTYPE


def _unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def pack(origins, dirs, tmax=100000.0, out=None):
    n = origins.shape[0]
    r = out if out is not None else np.zeros(n, dtype=RAY_DTYPE)
    r["ox"], r["oy"], r["oz"] = origins[:, 0], origins[:, 1], origins[:, 2]
    r["dx"], r["dy"], r["dz"] = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    r["tmin"] = 0.0
    r["tmax"] = tmax
    return r


def shell_rays(n, radius, seed=1, tmax=100000.0, out=None, chunk=1 << 22):
    """Incoherent outside-in stream; ~all rays hit a closed mesh of radius ~R centred at the origin."""
    rng = np.random.default_rng(seed)
    r = out if out is not None else np.zeros(n, dtype=RAY_DTYPE)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        o = _unit(rng.standard_normal((m, 3), dtype=np.float32)) * np.float32(3.0 * radius)
        tgt = _unit(rng.standard_normal((m, 3), dtype=np.float32)) * (np.float32(radius) * np.cbrt(rng.random((m, 1), dtype=np.float32)))
        pack(o, tgt - o, tmax, out=r[lo:lo + m])
    return r


def diffuse_bounce_rays(rays, hits, tris_u8, seed=2, tmax=100000.0):
    """One cosine-weighted bounce per hit about the interpolated shading normal (float32 numpy)."""
    rng = np.random.default_rng(seed)
    m = hits["tri"] != 0xFFFFFFFF
    idx = hits["tri"][m].astype(np.int64)
    f = tris_u8.view(np.float32).reshape(-1, 64)
    u, v = hits["u"][m, None], hits["v"][m, None]
    nrm = _unit(f[idx, 28:31] * u + f[idx, 48:51] * v + f[idx, 8:11] * (1 - u - v))
    o = np.stack([rays["ox"], rays["oy"], rays["oz"]], axis=1)[m]
    d = _unit(np.stack([rays["dx"], rays["dy"], rays["dz"]], axis=1)[m])
    pos = o + d * hits["t"][m, None]
    axis = np.where(np.abs(nrm[:, :1]) > 0.001, np.float32([[0, 1, 0]]), np.float32([[1, 0, 0]]))
    t = _unit(np.cross(axis, nrm))
    s = np.cross(nrm, t)
    phi = (2 * np.pi) * rng.random((idx.shape[0], 1), dtype=np.float32)
    r2 = rng.random((idx.shape[0], 1), dtype=np.float32)
    wi = _unit(s * np.cos(phi) * np.sqrt(r2) + t * np.sin(phi) * np.sqrt(r2) + nrm * np.sqrt(1 - r2)).astype(np.float32)
    return pack((pos + wi * np.float32(0.01)).astype(np.float32), wi, tmax)
