"""The reference's operator vocabulary (the MATLAB prototype's function names, matlab/*.m), served by the
CUDA library.  Argument order and meaning follow the .m files; arithmetic is the float32 CUDA path
(the bit-exact parity target), not MATLAB's double precision.

    F            = point_pair_feature(m_1, n_1, m_2, n_2)                      point_pair_feature.m:1-11
    F_disc       = my_discretize(F, d_dist, d_angle)                            my_discretize.m:3-4
    model, d_dist, d_angle = model_description(model_points, model_normals)     model_description.m:1-70
    result       = voting_scheme(model, model_points, model_normals,
                                 scene_points, scene_normals, d_dist, d_angle)  voting_scheme.m:1-150
    T_m_g, T_s_g, alpha = trans_model_scene(m_r, n_r_m, m_i, s_r, n_r_s, s_i)   trans_model_scene.m:1-41
"""
from __future__ import annotations

import numpy as np

from . import _capi as C
from .api import Model, Scene

D_ANGLE = np.float32(2.0) * np.float32(3.141592654) / np.float32(30)      # kernel.h:15-16, model_description.m:16-17


def _rows(a):
    a = np.ascontiguousarray(np.atleast_2d(np.asarray(a, np.float32)))
    if a.shape[1] != 3:
        raise ValueError("expected N x 3")
    return a


def point_pair_feature(m_1, n_1, m_2, n_2, d_dist: float = 1.0, return_keys: bool = False):
    """F = (|d|, angle(n1,d), angle(n2,d), angle(n1,n2)) for one pair or N pairs (row-wise)."""
    a, b, c, d = _rows(m_1), _rows(n_1), _rows(m_2), _rows(n_2)
    n = len(a)
    raw = np.empty((n, 4), np.float32)
    keys = np.empty(n, np.uint32)
    C.check(C.lib.ppf_point_pair_feature(a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data, n, float(d_dist),
                                         raw.ctypes.data, None, keys.ctypes.data if return_keys else None))
    out = raw[0] if np.ndim(m_1) == 1 else raw
    return (out, keys) if return_keys else out


def discretized_pair_feature(m_1, n_1, m_2, n_2, d_dist: float):
    """my_discretize(point_pair_feature(...)) and the table key, in one pass over N pairs."""
    a, b, c, d = _rows(m_1), _rows(n_1), _rows(m_2), _rows(n_2)
    n = len(a)
    disc = np.empty((n, 4), np.float32)
    keys = np.empty(n, np.uint32)
    C.check(C.lib.ppf_point_pair_feature(a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data, n, float(d_dist),
                                         None, disc.ctypes.data, keys.ctypes.data))
    return disc, keys


def my_discretize(F, d_dist, d_angle=D_ANGLE):
    """F - mod(F, step): exact restatement on the host (fmod is exact), float32 like the CUDA path."""
    F = np.asarray(F, np.float32)
    step = np.array([d_dist, d_angle, d_angle, d_angle], np.float32)
    return (F - np.fmod(F, step)).astype(np.float32)


def model_description(model_points, model_normals, d_dist: float | None = None, **model_kwargs):
    """Builds the PPF hash table.  d_dist defaults to the MATLAB rule 0.1 * max distance from the bounding-box
    centre (model_description.m:5-13); the CLI's rule is synth.d_dist_for (alignment.cpp:249-253)."""
    p = np.asarray(model_points, np.float32)
    if d_dist is None:
        centre = (p.min(0) + p.max(0)) / 2
        d_dist = float(np.float32(0.1) * np.float32(np.linalg.norm(p - centre, axis=1).max()))
    return Model(p, model_normals, d_dist, **model_kwargs), d_dist, float(D_ANGLE)


def voting_scheme(model: Model, model_points, model_normals, scene_points, scene_normals, d_dist=None, d_angle=None,
                  skip: int = 5):
    """Hough voting + pose recovery (voting_scheme.m; skip = reference-point stride, voting_scheme.m:10)."""
    scene = Scene(scene_points, scene_normals, model.d_dist if d_dist is None else d_dist, skip)
    return model.ppf_lookup(scene)


def trans_model_scene(m_r, n_r_m, m_i, s_r, n_r_s, s_i, return_index: bool = False):
    """T_m_g, T_s_g (4x4) and alpha (radians) for one tuple or N tuples (row-wise)."""
    arrs = [_rows(x) for x in (m_r, n_r_m, m_i, s_r, n_r_s, s_i)]
    n = len(arrs[0])
    Tm = np.empty((n, 4, 4), np.float32)
    Ts = np.empty((n, 4, 4), np.float32)
    al = np.empty(n, np.float32)
    ai = np.empty(n, np.uint32)
    C.check(C.lib.ppf_trans_model_scene(*[x.ctypes.data for x in arrs], n, Tm.ctypes.data, Ts.ctypes.data,
                                        al.ctypes.data, ai.ctypes.data))
    if np.ndim(m_r) == 1:
        Tm, Ts, al, ai = Tm[0], Ts[0], al[0], ai[0]
    return (Tm, Ts, al, ai) if return_index else (Tm, Ts, al)
