import numpy as np
import pytest

from objective_slam_b200 import io, synth


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("pcl_names", [False, True])
def test_ply_round_trip(tmp_path, fmt, pcl_names):
    p, n = synth.make_model(257, seed=3)
    path = tmp_path / "c.ply"
    io.write_ply(path, p, n, fmt=fmt, pcl_names=pcl_names)
    p2, n2 = io.read_ply(path)
    assert (p2.view(np.uint32) == p.view(np.uint32)).all() and (n2.view(np.uint32) == n.view(np.uint32)).all()


def test_ply_without_normals_and_with_faces(tmp_path):
    path = tmp_path / "mesh.ply"
    path.write_text("ply\nformat ascii 1.0\ncomment a mesh\nelement vertex 3\nproperty float x\nproperty float y\n"
                    "property float z\nproperty uchar red\nelement face 1\nproperty list uchar int vertex_indices\n"
                    "end_header\n0 0 0 255\n1 0 0 255\n0 1 0 255\n3 0 1 2\n")
    p, n = io.read_ply(path)
    assert n is None and p.shape == (3, 3) and p[1, 0] == 1


def test_pose_files_and_validation(tmp_path):
    mp, mn = synth.make_model(100, seed=1)
    sp, sn, T = synth.make_scene(mp, mn, 200, seed=2)
    io.write_pose(tmp_path / "gt.txt", T)
    assert np.allclose(io.read_pose(tmp_path / "gt.txt"), T, atol=1e-6)
    diam = io.model_diameter(mp)
    ok, dt, ang = io.validate_pose(T, T, diam)
    assert ok == 1 and dt == 0 and ang < 1e-6
    T2 = T.copy(); T2[:3, 3] += 0.2 * diam
    assert io.validate_pose(T2, T, diam)[0] == 0                      # translation gate: 0.1 x diameter
    c, s = np.cos(np.radians(13)), np.sin(np.radians(13))
    T3 = T.copy(); T3[:3, :3] = T[:3, :3] @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    ok, dt, ang = io.validate_pose(T3, T, diam)
    assert ok == 0 and abs(np.degrees(ang) - 13) < 1e-6               # rotation gate: 12 degrees


def test_voxel_grid_numpy_restatement_properties():
    from oracle.voxel_numpy import voxel_grid_downsample_numpy
    p, n = synth.make_model(5000, seed=4)
    q, m = voxel_grid_downsample_numpy(p, n, 10.0)
    assert 50 < len(q) < 1200
    cells = np.floor(q / 10.0).astype(int)
    assert len({tuple(c) for c in cells}) == len(q)                   # one point per leaf, centroid inside its leaf
    assert (np.linalg.norm(m, axis=1) <= 1.0 + 1e-5).all()            # averaged normals are not re-normalised
    assert np.linalg.norm(m, axis=1).min() < 0.999


def test_matlab_ply_write_fixture_and_trans_adj(tmp_path):
    """The layout matlab/utils/ply/ply_write.m emits for write_ply_cloud.m / compute_normals.m (committed fixture,
    tests/golden/make_ply_fixture.py) is read back to 6 decimals, our writer reproduces it byte for byte, and the
    .trans_adj side file round-trips (compute_normals.m:17-22)."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fix = os.path.join(here, "matlab_ply_write.ply")
    vals = np.load(os.path.join(here, "matlab_ply_write_values.npy"))
    p, n = io.read_ply(fix)
    assert p.shape == (12, 3) and np.abs(p - vals[:, :3]).max() < 1e-5 and np.abs(n - vals[:, 3:]).max() < 1e-6
    out = tmp_path / "again.ply"
    io.write_ply_matlab(out, vals[:, :3], vals[:, 3:])
    assert out.read_bytes() == open(fix, "rb").read()
    adj = io.read_trans_adj(fix)
    assert adj is not None and (p.min(0) > 0.99).all()                       # positive octant, min = 1 per axis
    assert np.allclose(adj, io.compute_trans_adj([vals[:, :3] - adj]), atol=1e-5)
    io.write_trans_adj(out, adj)
    assert open(str(out) + ".trans_adj").read() == open(fix + ".trans_adj").read()
    assert io.read_trans_adj(tmp_path / "missing.ply") is None
    # a pose between the original clouds, rewritten for the shifted ones, maps shifted model points onto shifted scene points
    mp, mn = synth.make_model(50, seed=1)
    sp, sn, T = synth.make_scene(mp, mn, 50, seed=2, noise=0.0, normal_noise_deg=0.0)
    a_m, a_s = np.array([3.0, 4.0, 5.0]), np.array([7.0, 1.0, 2.0])
    T2 = io.pose_in_adjusted_frame(T, a_m, a_s)
    x = io.apply_trans_adj(mp[:5], a_m).astype(np.float64)
    y = (T[:3, :3] @ mp[:5].astype(np.float64).T).T + T[:3, 3] + a_s
    assert np.abs((T2[:3, :3] @ x.T).T + T2[:3, 3] - y).max() < 1e-3
