"""GPU voxel grid against the numpy restatement of pcl::VoxelGrid (PCL is absent: parity unpinned beyond this)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,leaf", [(5000, 10.0), (200000, 2.5), (2, 0.5), (1000, 1000.0)])
def test_voxel_grid_matches_restatement(n, leaf):
    from objective_slam_b200 import synth
    from objective_slam_b200.voxel import voxel_grid_downsample
    from oracle.voxel_numpy import voxel_grid_downsample_numpy
    p, q = synth.make_model(n, seed=n)
    q = q * np.random.default_rng(1).uniform(0.5, 1.0, (n, 1)).astype(np.float32)
    a, b = voxel_grid_downsample(p, q, leaf)
    c, d = voxel_grid_downsample_numpy(p, q, leaf)
    assert a.shape == c.shape and len(a) >= 1
    assert np.abs(a - c).max() < 1e-3 and np.abs(b - d).max() < 1e-5     # float32 sums vs float64 sums
    if n == 1000:
        assert len(a) == 1 and np.allclose(a[0], p.mean(0), atol=1e-3)


def test_voxel_grid_device_tensors_and_nonfinite_points():
    import torch
    from objective_slam_b200 import synth
    from objective_slam_b200.voxel import voxel_grid_downsample
    p, q = synth.make_model(3000, seed=8)
    p[5] = np.nan; p[17, 1] = np.inf
    a, b = voxel_grid_downsample(p, q, 8.0)
    ta, tb = voxel_grid_downsample(torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda(), 8.0)
    assert np.isfinite(a).all() and (ta.cpu().numpy().view(np.uint32) == a.view(np.uint32)).all()
    assert (tb.cpu().numpy().view(np.uint32) == b.view(np.uint32)).all()


def test_downsampled_clouds_feed_the_hot_path():
    """alignment.cpp:265-298: scene downsampled with scene_leaf_size, model with its own d_dist, then registration."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import io, synth
    from objective_slam_b200.voxel import voxel_grid_downsample
    mp, mn = synth.make_model(20000, seed=21)
    sp, sn, T = synth.make_scene(mp, mn, 60000, seed=22)
    d = synth.d_dist_for(mp, 0.05)
    m_ds = voxel_grid_downsample(mp, mn, d)
    s_ds = voxel_grid_downsample(sp, sn, d)
    assert 500 < len(m_ds[0]) < 5000
    poses, status = ppf.ppf_registration([s_ds], [m_ds], [d], ref_point_downsample_factor=5)
    ok, dt, ang = io.validate_pose(poses[0, 0], T, io.model_diameter(mp))
    assert status[0, 0] == 0 and ok == 1, (dt, np.degrees(ang))
