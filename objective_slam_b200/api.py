"""Host-side mirror of the reference's operator surface for the PPF hot path.

Same names, argument meaning and error behaviour as the reference objects, so that
parity tests read like the reference's own call sites:

    Scene(points, normals, d_dist, ref_point_downsample_factor)     scene.h:14-16
    Model(points, normals, d_dist, vote_count_threshold, cpu_clustering,
          use_l1_norm, use_averaged_clusters)                       model.h:17-19
    model.ppf_lookup(scene)                                         model.h:35
    ppf_registration(scene_clouds, model_clouds, model_d_dists, ...) ppf.h:9-15

plus the MATLAB operator names of the prototype (point_pair_feature / my_discretize
-> ``Scene.features``; model_description -> ``Model.table``; voting_scheme ->
``Model.vote_histogram``; trans_model_scene -> ``LookupResult.transformations``).

Everything here is plumbing over the C ABI (include/ppf_b200.h); all arithmetic runs
in the CUDA library.  Clouds may be numpy arrays (host) or torch CUDA tensors (device).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import _capi as C


def _as_cloud_arrays(points, normals):
    """Return (xyz_ptr, xyz_stride, nrm_ptr, nrm_stride, n, mem, keepalive)."""
    try:
        import torch
    except Exception:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(points, torch.Tensor):
        if not (points.is_cuda and normals.is_cuda):
            points, normals = points.cpu().numpy(), normals.cpu().numpy()
        else:
            p = points.detach().to(torch.float32).contiguous()
            q = normals.detach().to(torch.float32).contiguous()
            if p.ndim != 2 or p.shape[1] != 3 or q.shape != p.shape:
                raise ValueError("clouds must be N x 3")
            torch.cuda.current_stream().synchronize()
            return p.data_ptr(), 3, q.data_ptr(), 3, p.shape[0], C.PPF_MEM_DEVICE, (p, q)
    p = np.ascontiguousarray(points, dtype=np.float32)
    q = np.ascontiguousarray(normals, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] != 3 or q.shape != p.shape:
        raise ValueError("clouds must be N x 3")
    # a valid pointer is needed even for n == 0
    if p.shape[0] == 0:
        p = np.zeros((1, 3), np.float32)[:0]
        q = np.zeros((1, 3), np.float32)[:0]
        buf = np.zeros(3, np.float32)
        return buf.ctypes.data, 3, buf.ctypes.data, 3, 0, C.PPF_MEM_HOST, (buf,)
    return p.ctypes.data, 3, q.ctypes.data, 3, p.shape[0], C.PPF_MEM_HOST, (p, q)


class Scene:
    """Scene::Scene (scene.cu:24-55): a cloud on the device plus the parameters the
    reference bakes into the scene's feature matrix (d_dist, reference-point stride)."""

    def __init__(self, points, normals, d_dist: float, ref_point_downsample_factor: int = 1):
        xp, xs, np_, ns, n, mem, keep = _as_cloud_arrays(points, normals)
        self._h = ctypes.c_void_p()
        C.check(C.lib.ppf_scene_create(xp, xs, np_, ns, n, mem, ctypes.byref(self._h)))
        self.d_dist = float(d_dist)
        self.ref_point_downsample_factor = int(ref_point_downsample_factor)
        self.n = n

    def numPoints(self) -> int:
        return C.lib.ppf_scene_num_points(self._h)

    def features(self, ref_range=None, other_range=None):
        """getModelPPFs / getHashKeys (scene.h:21-24) for a tile: (float32 [R,O,4], uint32 [R,O])."""
        rb, re = ref_range or (0, self.n)
        ob, oe = other_range or (0, self.n)
        ppf = np.empty((re - rb, oe - ob, 4), np.float32)
        keys = np.empty((re - rb, oe - ob), np.uint32)
        C.check(C.lib.ppf_scene_features(self._h, self.d_dist, self.ref_point_downsample_factor, rb, re, ob, oe,
                                         ppf.ctypes.data, keys.ctypes.data))
        return ppf, keys

    def close(self):
        if getattr(self, "_h", None) and C is not None and getattr(C, "lib", None) is not None:
            C.lib.ppf_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


@dataclass
class LookupResult:
    """Public members of Model after ppf_lookup (model.h:63-113) + the debug counters
    the reference logs (model.cu:122,152,161-168)."""
    votes: np.ndarray                 # uint64 [K]  [scene ref:32 | model pt:26 | alpha:6]
    voteCounts: np.ndarray            # uint32 [K]
    transformations: np.ndarray       # float32 [K,4,4]
    weightedVoteCounts: np.ndarray    # float32 [K]
    transformation_trans: np.ndarray  # float32 [K,3]
    transformation_rots: np.ndarray   # float32 [K,4]
    vote_counts_out: np.ndarray       # float32 [K]
    max_idx: int
    pose: np.ndarray                  # float32 [4,4]  (ppf.cu:80-93)
    num_nonunique_votes: int
    num_unique_votes: int
    num_top_votes: int
    max_vote_count: int
    num_scene_pairs: int
    num_exact_alpha: int
    ms_vote: float
    ms_finalize: float
    ms_pose_cluster: float
    status: int = C.PPF_OK


class Lookup:
    """Reusable device buffers of one Model::ppf_lookup."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        C.check(C.lib.ppf_lookup_create(ctypes.byref(self._h)))

    def stats(self) -> C.LookupStats:
        st = C.LookupStats()
        C.check(C.lib.ppf_lookup_get_stats(self._h, ctypes.byref(st)))
        return st

    def result(self, status=C.PPF_OK, arrays=True) -> LookupResult:
        st = self.stats()
        K = st.num_top_votes if arrays else 0
        votes = np.empty(K, np.uint64)
        counts = np.empty(K, np.uint32)
        T = np.empty((K, 4, 4), np.float32)
        w = np.empty(K, np.float32)
        tr = np.empty((K, 3), np.float32)
        rot = np.empty((K, 4), np.float32)
        sc = np.empty(K, np.float32)
        pose = np.zeros((4, 4), np.float32)
        if arrays:
            C.check(C.lib.ppf_lookup_get(self._h, votes.ctypes.data, counts.ctypes.data, T.ctypes.data, w.ctypes.data,
                                         tr.ctypes.data, rot.ctypes.data, sc.ctypes.data, pose.ctypes.data))
        else:
            C.check(C.lib.ppf_lookup_get(self._h, None, None, None, None, None, None, None, pose.ctypes.data))
        return LookupResult(votes, counts, T, w, tr, rot, sc, st.max_idx, pose, st.num_nonunique_votes,
                            st.num_unique_votes, st.num_top_votes, st.max_vote_count, st.num_scene_pairs,
                            st.num_exact_alpha, st.ms_vote, st.ms_finalize, st.ms_pose_cluster, status)

    def close(self):
        if getattr(self, "_h", None) and C is not None and getattr(C, "lib", None) is not None:
            C.lib.ppf_lookup_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Model:
    """Model::Model (model.cu:43-82): uploads the cloud and builds the PPF hash table."""

    def __init__(self, points, normals, d_dist: float, vote_count_threshold: float = 0.4,
                 cpu_clustering: bool = False, use_l1_norm: bool = False, use_averaged_clusters: bool = False,
                 expected_scene_points: int = 0):
        """expected_scene_points (optional, not in the reference): size of the scenes this model will meet;
        it only selects the table layout / vote kernel (ppf_set_expected_scene_points), never the results."""
        xp, xs, np_, ns, n, mem, keep = _as_cloud_arrays(points, normals)
        self._h = ctypes.c_void_p()
        self._lookup = None
        C.lib.ppf_set_expected_scene_points(int(expected_scene_points))
        try:
            C.check(C.lib.ppf_model_create(xp, xs, np_, ns, n, mem, float(d_dist), float(vote_count_threshold),
                                           int(use_l1_norm), int(use_averaged_clusters), ctypes.byref(self._h)))
        finally:
            C.lib.ppf_set_expected_scene_points(0)
        self.n = n
        self.d_dist = float(d_dist)
        self.vote_count_threshold = float(vote_count_threshold)
        self.cpu_clustering = bool(cpu_clustering)
        self.use_l1_norm = bool(use_l1_norm)
        self.use_averaged_clusters = bool(use_averaged_clusters)

    # -- persistent model database ------------------------------
    def save(self, path: str):
        """Write the built table to `path` (ppf_model_save)."""
        C.check(C.lib.ppf_model_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path: str, cpu_clustering: bool = False) -> "Model":
        """Model handle from a file written by save(): no rebuild (ppf_model_load)."""
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        self._lookup = None                         # before the load can fail: __del__ -> close() reads it
        C.check(C.lib.ppf_model_load(os.fsencode(path), ctypes.byref(self._h)))
        self.n = int(C.lib.ppf_model_num_points(self._h))
        d, thr, l1, avg = ctypes.c_float(), ctypes.c_float(), ctypes.c_int(), ctypes.c_int()
        C.check(C.lib.ppf_model_params(self._h, ctypes.byref(d), ctypes.byref(thr), ctypes.byref(l1), ctypes.byref(avg)))
        self.d_dist = float(d.value)
        self.vote_count_threshold = float(thr.value)
        self.use_l1_norm = bool(l1.value)
        self.use_averaged_clusters = bool(avg.value)
        self.cpu_clustering = bool(cpu_clustering)
        return self

    def layout(self):
        """(n_chunks, chunk_rows, grouped_kernel): how the table is laid out for voting."""
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        C.check(C.lib.ppf_model_layout(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, bool(c.value)

    # -- model_description -------------------------------------------------------
    def table(self):
        """ParallelHashArray contents: (hashkeys u32[U], counts u64[U], firstHashkeyIndex u64[U],
        hashkeyToDataMap u64[N*N])."""
        U, N = ctypes.c_size_t(), ctypes.c_size_t()
        C.check(C.lib.ppf_model_table_sizes(self._h, ctypes.byref(U), ctypes.byref(N)))
        hk = np.empty(U.value, np.uint32)
        cnt = np.empty(U.value, np.uint64)
        first = np.empty(U.value, np.uint64)
        mp = np.empty(N.value, np.uint64)
        C.check(C.lib.ppf_model_table_get(self._h, hk.ctypes.data, cnt.ctypes.data, first.ctypes.data, mp.ctypes.data))
        return hk, cnt, first, mp

    def features(self, ref_range=None, other_range=None):
        rb, re = ref_range or (0, self.n)
        ob, oe = other_range or (0, self.n)
        ppf = np.empty((re - rb, oe - ob, 4), np.float32)
        keys = np.empty((re - rb, oe - ob), np.uint32)
        C.check(C.lib.ppf_model_features(self._h, rb, re, ob, oe, ppf.ctypes.data, keys.ctypes.data))
        return ppf, keys

    # -- voting_scheme -------------------------------------------------------------
    def vote_histogram(self, scene: Scene, shard_rank: int = 0, shard_count: int = 1):
        """All unique vote codes and their counts (ascending code), before thresholding; optionally for one shard
        of the reference points only (reference point number shard_rank + k * shard_count)."""
        n = ctypes.c_size_t()
        df = scene.ref_point_downsample_factor
        C.check(C.lib.ppf_vote_histogram_shard(self._h, scene._h, df, shard_rank, shard_count, None, None, 0,
                                               ctypes.byref(n)))
        codes = np.empty(n.value, np.uint64)
        counts = np.empty(n.value, np.uint32)
        if n.value:
            C.check(C.lib.ppf_vote_histogram_shard(self._h, scene._h, df, shard_rank, shard_count, codes.ctypes.data,
                                                   counts.ctypes.data, n.value, ctypes.byref(n)))
        return codes, counts

    def ppf_lookup(self, scene: Scene, arrays: bool = True) -> LookupResult:
        """Model::ppf_lookup (model.cu:269-306)."""
        if self._lookup is None:
            self._lookup = Lookup()
        lk = self._lookup
        rc = C.check(C.lib.ppf_model_lookup(self._h, scene._h, scene.ref_point_downsample_factor, lk._h),
                     allow=(C.PPF_ERR_NO_VOTES,))
        res = lk.result(rc, arrays)
        if self.cpu_clustering and rc == C.PPF_OK:
            pose = np.zeros((4, 4), np.float32)
            C.check(C.lib.ppf_lookup_cluster_cpu(self._h, lk._h, pose.ctypes.data))
            res.pose = pose
        return res

    def close(self):
        if getattr(self, "_lookup", None) is not None:
            self._lookup.close()
            self._lookup = None
        if getattr(self, "_h", None) and C is not None and getattr(C, "lib", None) is not None:
            C.lib.ppf_model_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def ppf_registration(scene_clouds, model_clouds, model_d_dists, ref_point_downsample_factor=1,
                     vote_count_threshold=0.4, cpu_clustering=False, use_l1_norm=False,
                     use_averaged_clusters=False, devUse=0, model_weights=None, comm=None):
    """ppf_registration (ppf.h:9-15).  ``scene_clouds`` / ``model_clouds`` are lists of
    (points[N,3], normals[N,3]) host arrays; returns float32 [num_scenes, num_models, 4, 4] and the
    per-pair status codes.  With ``comm`` (dist.Comm) every rank makes the same call and the scene reference points
    are sharded over the ranks (ppf_registration_sharded): same poses on every rank."""
    keep = []

    def descs(clouds):
        arr = (C.CloudDesc * len(clouds))()
        for i, (p, q) in enumerate(clouds):
            p = np.ascontiguousarray(p, np.float32)
            q = np.ascontiguousarray(q, np.float32)
            keep.append((p, q))
            arr[i] = C.CloudDesc(p.ctypes.data, 3, q.ctypes.data, 3, len(p))
        return arr

    sd, md = descs(scene_clouds), descs(model_clouds)
    dd = np.ascontiguousarray(model_d_dists, np.float32)
    poses = np.zeros((len(scene_clouds), len(model_clouds), 4, 4), np.float32)
    status = np.zeros((len(scene_clouds), len(model_clouds)), np.int32)
    if comm is not None:
        C.check(C.lib.ppf_registration_sharded(sd, len(scene_clouds), md, len(model_clouds), dd.ctypes.data,
                                               int(ref_point_downsample_factor), float(vote_count_threshold),
                                               int(cpu_clustering), int(use_l1_norm), int(use_averaged_clusters),
                                               comm._h, poses.ctypes.data, status.ctypes.data))
        return poses, status
    C.check(C.lib.ppf_registration(sd, len(scene_clouds), md, len(model_clouds), dd.ctypes.data,
                                   int(ref_point_downsample_factor), float(vote_count_threshold),
                                   int(cpu_clustering), int(use_l1_norm), int(use_averaged_clusters), int(devUse),
                                   None, poses.ctypes.data, status.ctypes.data))
    return poses, status
