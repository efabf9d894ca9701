"""oracle/drost_m.c -- the MATLAB pipeline (drost.m) restated in C as a CPU timing baseline.  Parity is
unpinned (no MATLAB / Octave / JVM in the image), so the checks are the reference's own acceptance test
(alignment.cpp:317-323: translation within 0.1 diameter, rotation within 12 degrees) and determinism."""
import numpy as np

from objective_slam_b200 import synth
from oracle import cpu


def _angle(A, B):
    R = A[:3, :3].T @ B[:3, :3]
    return abs(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def test_matlab_pipeline_recovers_the_planted_pose():
    mp, mn = synth.make_model(300, seed=1)
    sp, sn, T = synth.make_scene(mp, mn, 600, seed=2)
    r = cpu.drost_m(mp, mn, sp, sn)                       # d_dist = 0.1 * max distance from the bbox centre
    assert r["pairs"] == 120 * 600 and r["votes"] > 0     # skip = 5 (voting_scheme.m:10)
    assert np.linalg.norm(r["pose"][:3, 3] - T[:3, 3]) < 0.1 * 100.0
    assert _angle(r["pose"], T) < np.radians(12)
    assert abs(np.linalg.det(r["pose"][:3, :3]) - 1) < 1e-9


def test_matlab_pipeline_is_deterministic_and_sampling_is_a_subset():
    mp, mn = synth.make_model(150, seed=3)
    sp, sn, _ = synth.make_scene(mp, mn, 300, seed=4)
    a = cpu.drost_m(mp, mn, sp, sn, threads=1)
    b = cpu.drost_m(mp, mn, sp, sn, threads=4)
    assert a["votes"] == b["votes"] and np.array_equal(a["pose"], b["pose"])
    c = cpu.drost_m(mp, mn, sp, sn, max_refs=10, scene_stride=3)
    assert 0 < c["votes"] < a["votes"] and c["pairs"] == 10 * 100
