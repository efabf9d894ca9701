"""ctypes bindings for lib/libppf_b200.so (the C ABI of include/ppf_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load,
importing this module raises.  Test oracles are never imported from here.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPF_B200_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("PPF_B200_LIB") or os.path.join(_HERE, "lib", "libppf_b200.so")

PPF_OK, PPF_ERR_INVALID, PPF_ERR_CUDA, PPF_ERR_UNSUPPORTED, PPF_ERR_NO_VOTES = range(5)
PPF_MEM_HOST, PPF_MEM_DEVICE = 0, 1

# every symbol include/ppf_b200.h declares (tests check that the library exports them all)
EXPORTS = [
    "ppf_last_error", "ppf_version", "ppf_kernel_launch_count", "ppf_release_cached_memory", "ppf_set_expected_scene_points",
    "ppf_current_stream",
    "ppf_scene_create", "ppf_scene_destroy", "ppf_scene_num_points", "ppf_scene_features",
    "ppf_model_create", "ppf_model_destroy", "ppf_model_num_points", "ppf_model_save", "ppf_model_load",
    "ppf_model_layout", "ppf_model_params", "ppf_model_table_sizes",
    "ppf_model_table_get", "ppf_model_features", "ppf_point_pair_feature", "ppf_trans_model_scene", "ppf_voxel_grid",
    "ppf_lookup_create", "ppf_lookup_destroy", "ppf_model_lookup", "ppf_lookup_vote",
    "ppf_lookup_local_max", "ppf_lookup_finalize", "ppf_lookup_survivors", "ppf_lookup_set_survivors",
    "ppf_lookup_copy_survivors",
    "ppf_lookup_poses", "ppf_lookup_cluster", "ppf_lookup_cluster_shard", "ppf_lookup_copy_scores",
    "ppf_lookup_set_scores", "ppf_lookup_cluster_finish", "ppf_lookup_cluster_cpu", "ppf_lookup_get_stats",
    "ppf_lookup_get", "ppf_vote_histogram", "ppf_vote_histogram_shard", "ppf_registration",
    "ppf_comm_unique_id", "ppf_comm_create_nccl", "ppf_comm_wrap_nccl", "ppf_comm_create_local", "ppf_comm_rank",
    "ppf_comm_size", "ppf_comm_destroy", "ppf_model_lookup_sharded", "ppf_registration_sharded",
]


class LookupStats(ctypes.Structure):
    _fields_ = [
        ("num_scene_pairs", ctypes.c_uint64),
        ("num_nonunique_votes", ctypes.c_uint64),
        ("num_unique_votes", ctypes.c_uint64),
        ("max_vote_count", ctypes.c_uint32),
        ("num_top_votes", ctypes.c_uint32),
        ("max_idx", ctypes.c_uint32),
        ("num_exact_alpha", ctypes.c_uint32),
        ("ms_vote", ctypes.c_float),
        ("ms_finalize", ctypes.c_float),
        ("ms_pose_cluster", ctypes.c_float),
    ]


class CloudDesc(ctypes.Structure):
    _fields_ = [
        ("xyz", ctypes.c_void_p), ("xyz_stride", ctypes.c_int),
        ("nrm", ctypes.c_void_p), ("nrm_stride", ctypes.c_int),
        ("n", ctypes.c_int),
    ]


class PpfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ppf_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C objective_slam_b200/csrc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint
    sz, u32 = ctypes.c_size_t, ctypes.c_uint32
    P = ctypes.POINTER
    L.ppf_last_error.restype = ctypes.c_char_p
    L.ppf_version.restype = ctypes.c_char_p
    L.ppf_kernel_launch_count.restype = ctypes.c_uint64
    L.ppf_release_cached_memory.restype = None
    L.ppf_set_expected_scene_points.argtypes = [ci]
    L.ppf_set_expected_scene_points.restype = None
    L.ppf_current_stream.argtypes = [P(vp)]
    L.ppf_scene_create.argtypes = [vp, ci, vp, ci, ci, ci, P(vp)]
    L.ppf_scene_destroy.argtypes = [vp]
    L.ppf_scene_destroy.restype = None
    L.ppf_scene_num_points.argtypes = [vp]
    L.ppf_scene_features.argtypes = [vp, cf, cu, ci, ci, ci, ci, vp, vp]
    L.ppf_model_create.argtypes = [vp, ci, vp, ci, ci, ci, cf, cf, ci, ci, P(vp)]
    L.ppf_model_destroy.argtypes = [vp]
    L.ppf_model_destroy.restype = None
    L.ppf_model_num_points.argtypes = [vp]
    L.ppf_model_save.argtypes = [vp, ctypes.c_char_p]
    L.ppf_model_load.argtypes = [ctypes.c_char_p, P(vp)]
    L.ppf_model_layout.argtypes = [vp, P(ci), P(ci), P(ci)]
    L.ppf_model_params.argtypes = [vp, P(cf), P(cf), P(ci), P(ci)]
    L.ppf_model_table_sizes.argtypes = [vp, P(sz), P(sz)]
    L.ppf_model_table_get.argtypes = [vp, vp, vp, vp, vp]
    L.ppf_model_features.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    L.ppf_voxel_grid.argtypes = [vp, ci, vp, ci, ci, ci, cf, vp, vp, P(ci)]
    L.ppf_point_pair_feature.argtypes = [vp, vp, vp, vp, sz, cf, vp, vp, vp]
    L.ppf_trans_model_scene.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.ppf_lookup_create.argtypes = [P(vp)]
    L.ppf_lookup_destroy.argtypes = [vp]
    L.ppf_lookup_destroy.restype = None
    L.ppf_model_lookup.argtypes = [vp, vp, cu, vp]
    L.ppf_lookup_vote.argtypes = [vp, vp, cu, ci, ci, vp]
    L.ppf_lookup_local_max.argtypes = [vp, P(u32)]
    L.ppf_lookup_finalize.argtypes = [vp, u32, vp]
    L.ppf_lookup_survivors.argtypes = [vp, P(sz), P(vp), P(vp)]
    L.ppf_lookup_set_survivors.argtypes = [vp, vp, vp, sz]
    L.ppf_lookup_copy_survivors.argtypes = [vp, vp, vp]
    L.ppf_lookup_poses.argtypes = [vp, vp, vp]
    L.ppf_lookup_cluster.argtypes = [vp, vp]
    L.ppf_lookup_cluster_shard.argtypes = [vp, vp, ci, ci]
    L.ppf_lookup_copy_scores.argtypes = [vp, vp]
    L.ppf_lookup_set_scores.argtypes = [vp, vp]
    L.ppf_lookup_cluster_finish.argtypes = [vp]
    L.ppf_lookup_cluster_cpu.argtypes = [vp, vp, vp]
    L.ppf_lookup_get_stats.argtypes = [vp, P(LookupStats)]
    L.ppf_lookup_get.argtypes = [vp] * 9
    L.ppf_vote_histogram.argtypes = [vp, vp, cu, vp, vp, sz, P(sz)]
    L.ppf_vote_histogram_shard.argtypes = [vp, vp, cu, ci, ci, vp, vp, sz, P(sz)]
    L.ppf_registration.argtypes = [P(CloudDesc), ci, P(CloudDesc), ci, vp, cu, cf, ci, ci, ci, ci, vp, vp, vp]
    L.ppf_comm_unique_id.argtypes = [vp]
    L.ppf_comm_create_nccl.argtypes = [vp, ci, ci, P(vp)]
    L.ppf_comm_wrap_nccl.argtypes = [vp, ci, ci, P(vp)]
    L.ppf_comm_create_local.argtypes = [ci, P(vp)]
    L.ppf_comm_rank.argtypes = [vp]
    L.ppf_comm_size.argtypes = [vp]
    L.ppf_comm_destroy.argtypes = [vp]
    L.ppf_comm_destroy.restype = None
    L.ppf_model_lookup_sharded.argtypes = [vp, vp, cu, vp, vp]
    L.ppf_registration_sharded.argtypes = [P(CloudDesc), ci, P(CloudDesc), ci, vp, cu, cf, ci, ci, ci, vp, vp, vp]
    return L


lib = _load()


def check(rc, allow=()):
    if rc != PPF_OK and rc not in allow:
        raise PpfError(rc, lib.ppf_last_error().decode())
    return rc
