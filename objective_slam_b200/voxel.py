"""voxelGridDownsample (alignment.cpp:79-87) on the GPU: one centroid per occupied leaf, positions and
(un-normalised) normals averaged, ascending leaf order (pcl::VoxelGrid semantics; PCL itself is absent, so
parity is checked against the numpy restatement below, which is test infrastructure)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi as C


def voxel_grid_downsample(points, normals, leaf: float):
    """GPU path: host numpy arrays or torch CUDA tensors in, same kind out."""
    try:
        import torch
    except Exception:  # pragma: no cover
        torch = None
    n_out = ctypes.c_int()
    if torch is not None and isinstance(points, torch.Tensor) and points.is_cuda:
        p = points.detach().to(torch.float32).contiguous()
        q = normals.detach().to(torch.float32).contiguous()
        op, oq = torch.empty_like(p), torch.empty_like(q)
        torch.cuda.current_stream().synchronize()
        C.check(C.lib.ppf_voxel_grid(p.data_ptr(), 3, q.data_ptr(), 3, p.shape[0], C.PPF_MEM_DEVICE, float(leaf),
                                     op.data_ptr(), oq.data_ptr(), ctypes.byref(n_out)))
        return op[: n_out.value], oq[: n_out.value]
    p = np.ascontiguousarray(points, np.float32)
    q = np.ascontiguousarray(normals, np.float32)
    op, oq = np.empty_like(p), np.empty_like(q)
    if len(p):
        C.check(C.lib.ppf_voxel_grid(p.ctypes.data, 3, q.ctypes.data, 3, len(p), C.PPF_MEM_HOST, float(leaf),
                                     op.ctypes.data, oq.ctypes.data, ctypes.byref(n_out)))
    return op[: n_out.value].copy(), oq[: n_out.value].copy()


def voxel_grid_downsample_numpy(points, normals, leaf: float):
    """Restatement of pcl::VoxelGrid<PointNormal>::applyFilter (checker for the tests; float64 sums)."""
    p = np.asarray(points, np.float32)
    q = np.asarray(normals, np.float32)
    ok = np.isfinite(p).all(1)
    p, q = p[ok], q[ok]
    if len(p) == 0:
        return p, q
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p * inv).astype(np.int64)
    min_b = np.floor(p.min(0) * inv).astype(np.int64)
    max_b = np.floor(p.max(0) * inv).astype(np.int64)
    div = max_b - min_b + 1
    cell = (ijk - min_b) @ np.array([1, div[0], div[0] * div[1]], np.int64)
    order = np.argsort(cell, kind="stable")
    cell_s = cell[order]
    heads = np.flatnonzero(np.r_[True, cell_s[1:] != cell_s[:-1]])
    counts = np.diff(np.r_[heads, len(cell_s)])
    sp = np.add.reduceat(p[order].astype(np.float64), heads) / counts[:, None]
    sq = np.add.reduceat(q[order].astype(np.float64), heads) / counts[:, None]
    return sp.astype(np.float32), sq.astype(np.float32)
