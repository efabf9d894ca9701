"""ctypes bindings for oracle/liboracle.so (ppf_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py may import this module; nothing under objective_slam_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


class OracleResult(ctypes.Structure):
    _fields_ = [
        ("num_scene_pairs", ctypes.c_uint64), ("num_nonunique_votes", ctypes.c_uint64),
        ("num_unique_votes", ctypes.c_uint64),
        ("max_vote_count", ctypes.c_uint32), ("K", ctypes.c_uint32), ("max_idx", ctypes.c_uint32),
        ("votes", ctypes.POINTER(ctypes.c_uint64)), ("counts", ctypes.POINTER(ctypes.c_uint32)),
        ("transformations", ctypes.POINTER(ctypes.c_float)), ("weighted", ctypes.POINTER(ctypes.c_float)),
        ("trans", ctypes.POINTER(ctypes.c_float)), ("rots", ctypes.POINTER(ctypes.c_float)),
        ("scores", ctypes.POINTER(ctypes.c_float)),
        ("pose", ctypes.c_float * 16),
        ("hist_n", ctypes.c_uint64),
        ("hist_codes", ctypes.POINTER(ctypes.c_uint64)), ("hist_counts", ctypes.POINTER(ctypes.c_uint32)),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, ci, cf, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint
        L.oracle_d_angle0.restype = cf
        L.oracle_hash.restype = ctypes.c_uint32
        L.oracle_hash.argtypes = [vp, ci]
        L.oracle_quant_downf.restype = cf
        L.oracle_quant_downf.argtypes = [cf, cf]
        L.oracle_disc_feature.argtypes = [vp, cf, cf, vp]
        L.oracle_rot.argtypes = [ci, cf, vp]
        L.oracle_trans_model_scene.restype = ctypes.c_uint32
        L.oracle_trans_model_scene.argtypes = [vp] * 6
        L.oracle_scene_features.argtypes = [vp, vp, ci, cf, cu, vp, vp]
        L.oracle_hash_array.argtypes = [vp, ctypes.c_size_t, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_lookup.argtypes = [vp, vp, ci, vp, vp, ci, cf, cu, cf, ci, ci, vp, vp, ci, ci, ci,
                                    ctypes.POINTER(OracleResult)]
        L.oracle_result_free.argtypes = [ctypes.POINTER(OracleResult)]
        L.oracle_time_voting.restype = ctypes.c_double
        L.oracle_time_voting.argtypes = [vp, vp, ci, vp, vp, ci, cf, cu, ci, ci, ci, ctypes.POINTER(ctypes.c_uint64),
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def scene_features(pts, nrm, d_dist, ref_df=1):
    pts, nrm = _f32(pts), _f32(nrm)
    n = len(pts)
    ppf = np.zeros((n, n, 4), np.float32)
    keys = np.zeros((n, n), np.uint32)
    lib().oracle_scene_features(_p(pts), _p(nrm), n, float(d_dist), int(ref_df), _p(ppf), _p(keys))
    return ppf, keys


def hash_array(keys):
    keys = np.ascontiguousarray(keys, np.uint32).ravel()
    n = len(keys)
    hk = np.zeros(max(n, 1), np.uint32)
    cnt = np.zeros(max(n, 1), np.uint64)
    first = np.zeros(max(n, 1), np.uint64)
    mp = np.zeros(max(n, 1), np.uint64)
    U = ctypes.c_size_t()
    lib().oracle_hash_array(_p(keys), n, _p(hk), _p(cnt), _p(first), _p(mp), ctypes.byref(U))
    return hk[:U.value].copy(), cnt[:U.value].copy(), first[:U.value].copy(), mp[:n].copy()


def lookup(mpts, mnrm, spts, snrm, d_dist, ref_df=1, thr=0.4, use_l1_norm=False, use_averaged_clusters=False,
           model_keys=None, scene_keys=None, histogram=False, per_vote_frames=False, threads=0):
    mpts, mnrm, spts, snrm = _f32(mpts), _f32(mnrm), _f32(spts), _f32(snrm)
    mk = np.ascontiguousarray(model_keys, np.uint32) if model_keys is not None else None
    sk = np.ascontiguousarray(scene_keys, np.uint32) if scene_keys is not None else None
    r = OracleResult()
    rc = lib().oracle_lookup(_p(mpts), _p(mnrm), len(mpts), _p(spts), _p(snrm), len(spts), float(d_dist), int(ref_df),
                             float(thr), int(use_l1_norm), int(use_averaged_clusters), _p(mk), _p(sk), int(histogram),
                             int(per_vote_frames), int(threads), ctypes.byref(r))
    if rc:
        raise ValueError("oracle_lookup failed")
    K = r.K

    def arr(ptr, n, dt):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)

    out = dict(
        K=K, num_scene_pairs=r.num_scene_pairs, num_nonunique_votes=r.num_nonunique_votes,
        num_unique_votes=r.num_unique_votes, max_vote_count=r.max_vote_count, max_idx=r.max_idx,
        votes=arr(r.votes, K, np.uint64), counts=arr(r.counts, K, np.uint32),
        transformations=arr(r.transformations, 16 * K, np.float32).reshape(K, 4, 4),
        weighted=arr(r.weighted, K, np.float32), trans=arr(r.trans, 3 * K, np.float32).reshape(K, 3),
        rots=arr(r.rots, 4 * K, np.float32).reshape(K, 4), scores=arr(r.scores, K, np.float32),
        pose=np.array(list(r.pose), np.float32).reshape(4, 4),
    )
    if histogram:
        out["hist_codes"] = arr(r.hist_codes, r.hist_n, np.uint64)
        out["hist_counts"] = arr(r.hist_counts, r.hist_n, np.uint32)
    lib().oracle_result_free(ctypes.byref(r))
    return out


def drost_m(mpts, mnrm, spts, snrm, d_dist=0.0, skip=5, max_refs=0, threads=0, scene_stride=1):
    """oracle/drost_m.c: the MATLAB pipeline (model_description.m + voting_scheme.m, double precision).
    Returns dict(seconds, build_seconds, pairs, votes, pose[4,4] float64).  Timing baseline, parity unpinned."""
    L = lib()
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.drost_m_run.restype = cd
    L.drost_m_run.argtypes = [vp, vp, ci, vp, vp, ci, cd, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_uint64),
                              ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(cd), vp]
    mpts, mnrm, spts, snrm = _f32(mpts), _f32(mnrm), _f32(spts), _f32(snrm)
    pairs, votes, build_s = ctypes.c_uint64(), ctypes.c_uint64(), cd()
    pose = np.zeros((4, 4), np.float64)
    s = L.drost_m_run(_p(mpts), _p(mnrm), len(mpts), _p(spts), _p(snrm), len(spts), float(d_dist), int(skip),
                      int(max_refs), int(scene_stride), int(threads), ctypes.byref(pairs), ctypes.byref(votes),
                      ctypes.byref(build_s), pose.ctypes.data)
    return dict(seconds=s, build_seconds=build_s.value, pairs=pairs.value, votes=votes.value, pose=pose)


def pcl_style(mpts, mnrm, spts, snrm, d_dist, ref_rate=5, max_refs=0, threads=0, scene_stride=1):
    """oracle/pcl_style.c: Drost's registration organised like PCL's PPFRegistration (alpha_m precomputed, one
    alpha_s per scene pair, per-reference accumulator, greedy pose clustering).  Timing baseline, parity unpinned.
    Returns dict(seconds, build_seconds, pairs, votes, pose[4,4] float64)."""
    L = lib()
    vp, ci, cd, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_float
    L.pcl_style_run.restype = cd
    L.pcl_style_run.argtypes = [vp, vp, ci, vp, vp, ci, cf, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_uint64),
                                ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(cd), vp]
    mpts, mnrm, spts, snrm = _f32(mpts), _f32(mnrm), _f32(spts), _f32(snrm)
    pairs, votes, build_s = ctypes.c_uint64(), ctypes.c_uint64(), cd()
    pose = np.zeros((4, 4), np.float64)
    s = L.pcl_style_run(_p(mpts), _p(mnrm), len(mpts), _p(spts), _p(snrm), len(spts), float(d_dist), int(ref_rate),
                        int(max_refs), int(scene_stride), int(threads), ctypes.byref(pairs), ctypes.byref(votes),
                        ctypes.byref(build_s), pose.ctypes.data)
    if s < 0:
        raise ValueError("pcl_style: models of >= 65536 points are not supported")
    return dict(seconds=s, build_seconds=build_s.value, pairs=pairs.value, votes=votes.value, pose=pose)


def time_voting(mpts, mnrm, spts, snrm, d_dist, ref_df=1, max_refs=0, threads=0, scene_stride=1):
    """Seconds spent voting over (up to max_refs) reference points; returns dict."""
    mpts, mnrm, spts, snrm = _f32(mpts), _f32(mnrm), _f32(spts), _f32(snrm)
    pairs, votes, build_s = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_double()
    s = lib().oracle_time_voting(_p(mpts), _p(mnrm), len(mpts), _p(spts), _p(snrm), len(spts), float(d_dist),
                                 int(ref_df), int(max_refs), int(scene_stride), int(threads), ctypes.byref(pairs), ctypes.byref(votes),
                                 ctypes.byref(build_s))
    return dict(seconds=s, pairs=pairs.value, votes=votes.value, build_seconds=build_s.value)
