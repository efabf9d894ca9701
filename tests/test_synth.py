import numpy as np

from objective_slam_b200 import synth


def test_model_is_deterministic_positive_octant_unit_normals():
    p, n = synth.make_model(500, seed=7)
    p2, n2 = synth.make_model(500, seed=7)
    assert (p == p2).all() and (n == n2).all()
    assert p.dtype == np.float32 and p.shape == (500, 3)
    assert p.min() >= 1.0 - 1e-4                                   # reference needs the positive octant
    assert abs((p.max(0) - p.min(0)).max() - 100.0) < 1e-3
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    c = p.mean(0)
    assert ((p - c) * n).sum(1).min() > 0                          # outward normals


def test_scene_contains_the_transformed_model():
    mp, mn = synth.make_model(300, seed=3)
    sp, sn, T = synth.make_scene(mp, mn, 700, seed=4, noise=0.0, normal_noise_deg=0.0, shrink_normals=False)
    assert sp.shape == (700, 3) and sp.min() >= 1.0 - 1e-3
    moved = mp.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    d = np.linalg.norm(moved[:, None, :] - sp[None, :, :].astype(np.float64), axis=2).min(1)
    assert d.max() < 1e-3                                          # every model point appears in the scene
    assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9


def test_d_dist_rule():
    mp, _ = synth.make_model(100)
    assert abs(synth.d_dist_for(mp, 0.05) - 5.0) < 1e-3            # alignment.cpp:249-253


def test_lattice_scene_has_axis_aligned_normals():
    p, n = synth.make_lattice_scene(3000)
    assert len(p) == 3000 and set(np.unique(n)) <= {0.0, 1.0}
