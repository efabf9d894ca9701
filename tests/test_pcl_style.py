"""oracle/pcl_style.c -- Drost's registration organised like PCL's PPFRegistration (alpha_m precomputed, one
alpha_s per scene pair, per-reference accumulator, greedy pose clustering), the third CPU timing baseline.
Parity is unpinned (PCL is not installed), so the checks are the reference's own acceptance test
(alignment.cpp:317-323) and determinism."""
import numpy as np

from objective_slam_b200 import synth
from oracle import cpu


def _angle(A, B):
    R = A[:3, :3].T @ B[:3, :3]
    return abs(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def test_pcl_style_registration_recovers_the_planted_pose():
    for nm, ns, s1, s2 in ((300, 600, 1, 2), (500, 1500, 3, 4)):
        mp, mn = synth.make_model(nm, seed=s1)
        sp, sn, T = synth.make_scene(mp, mn, ns, seed=s2)
        r = cpu.pcl_style(mp, mn, sp, sn, synth.d_dist_for(mp))
        assert r["pairs"] == ((ns + 4) // 5) * ns and r["votes"] > 0
        assert np.linalg.norm(r["pose"][:3, 3] - T[:3, 3]) < 0.1 * 100.0
        assert _angle(r["pose"], T) < np.radians(12)
        assert abs(np.linalg.det(r["pose"][:3, :3]) - 1) < 1e-5


def test_pcl_style_is_deterministic_and_sampling_is_a_subset():
    mp, mn = synth.make_model(200, seed=5)
    sp, sn, _ = synth.make_scene(mp, mn, 400, seed=6)
    d = synth.d_dist_for(mp)
    a = cpu.pcl_style(mp, mn, sp, sn, d, threads=1)
    b = cpu.pcl_style(mp, mn, sp, sn, d, threads=4)
    assert a["votes"] == b["votes"] and np.array_equal(a["pose"], b["pose"])
    c = cpu.pcl_style(mp, mn, sp, sn, d, max_refs=10, scene_stride=4)
    assert 0 < c["votes"] < a["votes"] and c["pairs"] == 10 * 100
