"""pcl::VoxelGrid<PointNormal>::applyFilter restated in numpy -- TEST INFRASTRUCTURE ONLY (checker of
tests/test_voxel_gpu.py and tests/test_io.py for csrc/ppf_voxel.cu; nothing under objective_slam_b200/ imports it).

PARITY UNPINNED: PCL (pcl/filters/voxel_grid.h, the library alignment.cpp:79-87 calls) is not vendored in
/root/reference and not installed, so this follows the published algorithm: leaf index = floor(p / leaf) relative to
the minimum leaf of the cloud, points grouped by linearised leaf index (x fastest), one output point per occupied
leaf = centroid of positions and of the (un-normalised) normals, leaves in ascending index order; non-finite points
are dropped (the filter's default for dense = false clouds)."""
import numpy as np


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
