"""Deterministic synthetic clouds for parity tests and bench.py.

The reference ships no data (its .gitignore drops *.ply / *.pcd), so every input
is generated here.  The shapes follow what the reference's own pipeline feeds
into ``ppf_registration`` (alignment.cpp:246-298):

* a model sampled on a closed "lumpy sphere" with analytic outward normals,
  scaled to a diameter of 100 units and shifted into the positive octant (the
  reference requires it: scene_generation.hpp:95-96, compute_trans_adj.m:8-10);
* a scene = model under a known rigid transform (Shoemake uniform quaternion,
  scene_generation.hpp:33-51) + Gaussian noise + planar clutter, normals left
  un-normalised (PCL's VoxelGrid averages them, alignment.cpp:79-87);
* ``d_dist = tau_d * max(bbox extent)`` (alignment.cpp:249-253).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xD205


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def shoemake_rotation(rng) -> np.ndarray:
    """Uniform random rotation (same construction as scene_generation.hpp:33-51)."""
    u1, u2, u3 = rng.random(3)
    q = np.array([
        np.sqrt(1 - u1) * np.sin(2 * np.pi * u2),
        np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
        np.sqrt(u1) * np.sin(2 * np.pi * u3),
        np.sqrt(u1) * np.cos(2 * np.pi * u3),
    ])
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def make_model(n_points: int, seed: int = SEED_BASE, diameter: float = 100.0):
    """Area-weighted samples of the star-shaped surface r(u) = 1 + sum_k a_k sin(f_k c_k.u + phi_k).

    Ten bumps of amplitude 0.05-0.10 and angular frequency 2-7 make the normals swing far from
    radial, so the quantised pair features spread over thousands of bins the way a real scanned
    object's do (a near-sphere would collapse them onto a one-parameter family).
    Returns (points[N,3] float32, normals[N,3] float32), points in the positive octant with min
    coordinate 1, normals unit length and outward (analytic).
    """
    rng = np.random.default_rng(seed)
    K = 10
    c = _unit(rng.normal(size=(K, 3)))
    f = rng.integers(2, 8, size=K).astype(np.float64)
    a = rng.uniform(0.05, 0.10, size=K)
    phi = rng.uniform(0, 2 * np.pi, size=K)

    def surface(u):
        arg = (u @ c.T) * f + phi
        r = 1.0 + (a * np.sin(arg)).sum(axis=1)
        g = ((a * f * np.cos(arg))[:, :, None] * c[None, :, :]).sum(axis=1)   # grad_u r
        gt = g - (g * u).sum(axis=1, keepdims=True) * u                        # tangential part
        n = _unit(u - gt / r[:, None])
        return r, n

    pts, nrm = [], []
    need = n_points
    while need > 0:
        u = _unit(rng.normal(size=(8 * need + 64, 3)))
        r, n = surface(u)
        w = r * r / np.maximum((n * u).sum(axis=1), 0.1)     # dA / dOmega
        keep = rng.random(len(w)) < w / (2.0 ** 2 / 0.1)
        u, r, n = u[keep][:need], r[keep][:need], n[keep][:need]
        pts.append(u * r[:, None])
        nrm.append(n)
        need -= len(u)
    p = np.concatenate(pts)
    n = np.concatenate(nrm)
    ext = (p.max(axis=0) - p.min(axis=0)).max()
    p = p * (diameter / ext)
    p = p - p.min(axis=0) + 1.0
    return p.astype(np.float32), n.astype(np.float32)


def d_dist_for(points: np.ndarray, tau_d: float = 0.05) -> float:
    """alignment.cpp:249-253: tau_d * max extent of the axis-aligned bounding box."""
    p = np.asarray(points, np.float32)
    ext = (p.max(axis=0) - p.min(axis=0)).max()
    return float(np.float32(tau_d) * np.float32(ext))


def make_scene(model_pts, model_nrm, n_points: int, seed: int = SEED_BASE + 1,
               noise: float = 0.005, normal_noise_deg: float = 2.0, box: float = 3.0,
               shrink_normals: bool = True):
    """Scene = transformed noisy copy of the model + planar clutter.

    Returns (points, normals, T) where T (4x4 float64) maps model -> scene
    coordinates (the ground truth the reference validates against,
    alignment.cpp:300-335).  If ``n_points`` is smaller than the model, a random
    subset of the model is used and no clutter is added.
    """
    rng = np.random.default_rng(seed)
    mp = np.asarray(model_pts, np.float64)
    mn = np.asarray(model_nrm, np.float64)
    diam = (mp.max(axis=0) - mp.min(axis=0)).max()
    R = shoemake_rotation(rng)
    centre = mp.mean(axis=0)
    t = rng.random(3) * diam
    n_obj = min(len(mp), n_points)
    sel = np.sort(rng.choice(len(mp), n_obj, replace=False)) if n_obj < len(mp) else np.arange(len(mp))
    op = (mp[sel] - centre) @ R.T + centre + t
    on = mn[sel] @ R.T
    op = op + rng.normal(scale=noise * diam, size=op.shape)
    on = _unit(on + rng.normal(scale=np.tan(np.radians(normal_noise_deg)), size=on.shape))

    n_clut = n_points - n_obj
    cp = np.zeros((0, 3))
    cn = np.zeros((0, 3))
    if n_clut > 0:
        n_planes = 6
        per = np.full(n_planes, n_clut // n_planes)
        per[: n_clut - per.sum()] += 1
        lo = centre + t - 0.5 * box * diam
        ps, ns = [], []
        for k in range(n_planes):
            nk = _unit(rng.normal(size=3))
            e1 = _unit(np.cross(nk, rng.normal(size=3)))
            e2 = np.cross(nk, e1)
            origin = lo + rng.random(3) * box * diam
            side = rng.uniform(1.0, 2.0) * diam
            uv = (rng.random((per[k], 2)) - 0.5) * side
            q = origin + uv[:, :1] * e1 + uv[:, 1:] * e2
            q = q + nk * rng.normal(scale=noise * diam, size=(per[k], 1))
            ps.append(q)
            ns.append(_unit(nk + rng.normal(scale=np.tan(np.radians(normal_noise_deg)), size=(per[k], 3))))
        cp = np.concatenate(ps)
        cn = np.concatenate(ns)
    p = np.concatenate([op, cp])
    n = np.concatenate([on, cn])
    if shrink_normals:
        n = n * rng.uniform(0.8, 1.0, size=(len(n), 1))
    perm = rng.permutation(len(p))
    p, n = p[perm], n[perm]
    shift = 1.0 - p.min(axis=0)
    p = p + shift
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = centre + t + shift - R @ centre
    return p.astype(np.float32), n.astype(np.float32), T


def make_lattice_scene(n_points: int, pitch: float = 0.5, seed: int = SEED_BASE + 5):
    """TSDF-like cloud (BASELINE config 5): points on a regular lattice of an
    axis-aligned room (floor + two walls), axis-aligned normals.  Exercises the
    exact-zero / signed-zero corners of trans_model_scene."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n_points / 3.0)))
    g = np.arange(side) * pitch + 1.0
    a, b = np.meshgrid(g, g, indexing="ij")
    a, b = a.ravel(), b.ravel()
    one = np.ones_like(a)
    floor = np.stack([a, b, one], 1); fn = np.tile([0.0, 0.0, 1.0], (len(a), 1))
    wall1 = np.stack([a, one, b], 1); w1n = np.tile([0.0, 1.0, 0.0], (len(a), 1))
    wall2 = np.stack([one, a, b], 1); w2n = np.tile([1.0, 0.0, 0.0], (len(a), 1))
    p = np.concatenate([floor, wall1, wall2])
    n = np.concatenate([fn, w1n, w2n])
    perm = rng.permutation(len(p))[:n_points]
    return p[perm].astype(np.float32), n[perm].astype(np.float32)
