"""ctypes bindings for oracle/_ref/libppf_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own kernel.cu / vector_ops.cu /
parallel_hash_array.hpp compiled for sm_100a (oracle/Makefile) plus the replay
harness oracle/ref_harness.cu.  Only tests/, __graft_entry__.smoke() and the
reference / cpu_baseline legs of bench.py may import this module.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libppf_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_LIB_PATH)
        vp, ci, cf, cu, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_long
        L.ref_scene_create.restype = vp
        L.ref_scene_create.argtypes = [vp, vp, ci, cf, cu]
        L.ref_scene_destroy.argtypes = [vp]
        L.ref_scene_get.argtypes = [vp, vp, vp]
        L.ref_model_create.restype = vp
        L.ref_model_create.argtypes = [vp, vp, ci, cf]
        L.ref_model_destroy.argtypes = [vp]
        L.ref_model_table_sizes.argtypes = [vp, vp]
        L.ref_model_table_get.argtypes = [vp, vp, vp, vp, vp]
        L.ref_model_get.argtypes = [vp, vp]
        L.ref_ppf_lookup.restype = cl
        L.ref_ppf_lookup.argtypes = [vp, vp, cf, ci, ci]
        L.ref_lookup_stats.argtypes = [vp, vp]
        L.ref_lookup_get.argtypes = [vp] * 9
        L.ref_vote_histogram.restype = cl
        L.ref_vote_histogram.argtypes = [vp, vp, vp, vp, cl]
        L.ref_time_scene_lookup.restype = cf
        L.ref_time_scene_lookup.argtypes = [vp, vp, vp, ci, cu, cf]
        L.ref_modeb_histogram.restype = cl
        L.ref_modeb_histogram.argtypes = [vp, vp, vp, ci, vp, ci, vp, vp, cl, vp]
        # host shims
        L.refhost_hash.restype = cu
        L.refhost_hash.argtypes = [vp, ci]
        L.refhost_quant_downf.restype = cf
        L.refhost_quant_downf.argtypes = [cf, cf]
        L.refhost_d_angle0.restype = cf
        L.refhost_dot3.restype = cf
        L.refhost_dot3.argtypes = [vp, vp]
        L.refhost_rot.argtypes = [ci, cf, vp]
        L.refhost_disc_feature.argtypes = [vp, cf, cf, vp]
        L.refhost_discretize.argtypes = [vp, cf, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefScene:
    """Replays Scene::Scene (scene.cu:24-55)."""

    def __init__(self, pts, nrm, d_dist, ref_df=1):
        self.pts, self.nrm = _f32(pts), _f32(nrm)
        self.n = len(self.pts)
        self.h = lib().ref_scene_create(_p(self.pts), _p(self.nrm), self.n, float(d_dist), int(ref_df))

    def features(self):
        ppf = np.empty((self.n, self.n, 4), np.float32)
        keys = np.empty((self.n, self.n), np.uint32)
        lib().ref_scene_get(self.h, _p(ppf), _p(keys))
        return ppf, keys

    def close(self):
        if self.h:
            lib().ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefModel:
    """Replays Model::Model (model.cu:43-82) and Model::ppf_lookup (model.cu:269-306)."""

    def __init__(self, pts, nrm, d_dist):
        self.pts, self.nrm = _f32(pts), _f32(nrm)
        self.n = len(self.pts)
        self.d_dist = float(d_dist)
        self.h = lib().ref_model_create(_p(self.pts), _p(self.nrm), self.n, self.d_dist)

    def table(self):
        sizes = np.zeros(2, np.uint64)
        lib().ref_model_table_sizes(self.h, _p(sizes))
        U, N = int(sizes[0]), int(sizes[1])
        hk = np.empty(U, np.uint32)
        cnt = np.empty(U, np.uint64)
        first = np.empty(U, np.uint64)
        mp = np.empty(N, np.uint64)
        lib().ref_model_table_get(self.h, _p(hk), _p(cnt), _p(first), _p(mp))
        return hk, cnt, first, mp

    def features(self):
        ppf = np.empty((self.n, self.n, 4), np.float32)
        lib().ref_model_get(self.h, _p(ppf))
        return ppf

    def lookup(self, scene: RefScene, vote_count_threshold=0.4, use_l1_norm=False,
               use_averaged_clusters=False):
        K = lib().ref_ppf_lookup(self.h, scene.h, float(vote_count_threshold), int(use_l1_norm),
                                 int(use_averaged_clusters))
        stats = np.zeros(4, np.uint64)
        lib().ref_lookup_stats(self.h, _p(stats))
        out = dict(K=int(K), num_nonunique_votes=int(stats[0]), num_unique_votes=int(stats[1]),
                   max_idx=int(stats[3]))
        K = max(int(K), 0)
        votes = np.empty(K, np.uint64)
        counts = np.empty(K, np.uint32)
        T = np.empty((K, 4, 4), np.float32)
        w = np.empty(K, np.float32)
        tr = np.empty((K, 3), np.float32)
        rot = np.empty((K, 4), np.float32)
        sc = np.empty(K, np.float32)
        pose = np.zeros((4, 4), np.float32)
        lib().ref_lookup_get(self.h, _p(votes), _p(counts), _p(T), _p(w), _p(tr), _p(rot), _p(sc), _p(pose))
        out.update(votes=votes, counts=counts, transformations=T, weighted=w, trans=tr, rots=rot,
                   scores=sc, pose=pose)
        return out

    def vote_histogram(self, scene: RefScene):
        n = lib().ref_vote_histogram(self.h, scene.h, None, None, 0)
        codes = np.empty(n, np.uint64)
        counts = np.empty(n, np.uint32)
        if n:
            lib().ref_vote_histogram(self.h, scene.h, _p(codes), _p(counts), n)
        return codes, counts

    def modeb_histogram(self, pts, nrm, refs):
        """Oracle Mode B: every non-zero accumulator cell (code, count; ascending code) of the scene
        reference points `refs` (indices into the scene cloud), computed by the reference's device functions
        streamed over the scene -- for sizes the whole-kernel replay (Mode A) cannot hold.  Also returns the
        number of votes cast."""
        pts, nrm = _f32(pts), _f32(nrm)
        refs = np.ascontiguousarray(refs, np.int32)
        nv = np.zeros(1, np.uint64)
        n = lib().ref_modeb_histogram(self.h, _p(pts), _p(nrm), len(pts), _p(refs), len(refs), None, None, 0, _p(nv))
        if n < 0:
            raise RuntimeError("ref_modeb_histogram: CUDA error")
        codes = np.empty(n, np.uint64)
        counts = np.empty(n, np.uint32)
        if n:
            lib().ref_modeb_histogram(self.h, _p(pts), _p(nrm), len(pts), _p(refs), len(refs), _p(codes), _p(counts), n, _p(nv))
        return codes, counts, int(nv[0])

    def time_scene_lookup(self, pts, nrm, ref_df=1, thr=0.4):
        pts, nrm = _f32(pts), _f32(nrm)
        return lib().ref_time_scene_lookup(self.h, _p(pts), _p(nrm), len(pts), int(ref_df), float(thr))

    def close(self):
        if self.h:
            lib().ref_model_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
