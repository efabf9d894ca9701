"""The MATLAB-named operator surface (objective_slam_b200.operators) against the CPU oracle and the reference
kernels: point_pair_feature / my_discretize / trans_model_scene / model_description / voting_scheme."""
import ctypes

import numpy as np
import pytest

from conftest import have_ref

pytestmark = pytest.mark.gpu


def _pairs(n, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda s: (rng.normal(size=(n, 3)) * s).astype(np.float32)
    return f(40) + 60, f(1), f(40) + 60, f(1)


def test_point_pair_feature_and_discretize_match_the_scene_kernel():
    import objective_slam_b200 as ppf
    from objective_slam_b200 import operators as ops, synth
    mp, mn = synth.make_model(120, seed=5)
    d = synth.d_dist_for(mp)
    ppfs, keys = ppf.Scene(mp, mn, d, 1).features()
    i, j = np.meshgrid(np.arange(120), np.arange(120), indexing="ij")
    sel = i != j
    disc, k = ops.discretized_pair_feature(mp[i[sel]], mn[i[sel]], mp[j[sel]], mn[j[sel]], d)
    assert (disc.view(np.uint32) == ppfs[sel].view(np.uint32)).all() and (k == keys[sel]).all()
    raw = ops.point_pair_feature(mp[i[sel]], mn[i[sel]], mp[j[sel]], mn[j[sel]])
    # my_discretize on the host (exact fmod) of the GPU's raw features == the GPU's closed-form quantiser
    assert (ops.my_discretize(raw, d).view(np.uint32) == disc.view(np.uint32)).all()
    one = ops.point_pair_feature(mp[0], mn[0], mp[1], mn[1])
    assert one.shape == (4,) and (one.view(np.uint32) == raw[0].view(np.uint32)).all()


def test_point_pair_feature_close_to_cpu_oracle():
    from objective_slam_b200 import operators as ops
    from oracle import cpu
    p1, n1, p2, n2 = _pairs(2000, 1)
    F = ops.point_pair_feature(p1, n1, p2, n2)
    L = cpu.lib()
    ref = np.zeros(4, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for t in range(0, 2000, 37):
        L.oracle_compute_ppf(P(p1[t]), P(n1[t]), P(p2[t]), P(n2[t]), P(ref))
        assert np.allclose(F[t], ref, rtol=2e-6, atol=2e-6)


def test_trans_model_scene_matches_oracle_and_is_a_rigid_frame():
    from objective_slam_b200 import operators as ops
    from oracle import cpu
    rng = np.random.default_rng(3)
    n = 500
    v = lambda s, o=0: (rng.normal(size=(n, 3)) * s + o).astype(np.float32)
    m_r, n_r_m, m_i, s_r, n_r_s, s_i = v(30, 50), v(1), v(30, 50), v(30, 50), v(1), v(30, 50)
    Tm, Ts, al, ai = ops.trans_model_scene(m_r, n_r_m, m_i, s_r, n_r_s, s_i, return_index=True)
    # T_g maps the reference point to the origin and its normal onto +x (trans_model_scene.m)
    for T, p, nrm in ((Tm, m_r, n_r_m), (Ts, s_r, n_r_s)):
        R = T[:, :3, :3].astype(np.float64)
        assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-5
        o = np.einsum("nij,nj->ni", T[:, :3, :3], p) + T[:, :3, 3]
        assert np.abs(o).max() < 2e-3
        nx = np.einsum("nij,nj->ni", T[:, :3, :3], nrm / np.linalg.norm(nrm, axis=1, keepdims=True))
        assert np.abs(nx[:, 0] - 1).max() < 1e-5
    L = cpu.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    same = sum(L.oracle_trans_model_scene(P(m_r[t]), P(n_r_m[t]), P(m_i[t]), P(s_r[t]), P(n_r_s[t]), P(s_i[t])) == ai[t]
               for t in range(n))
    assert same >= n - 2                                       # libm vs GPU atan2f may flip a bin edge
    assert ((al >= -np.pi - 1e-6) & (al <= np.pi + 1e-6)).all() and (ai <= 30).all()
    assert (np.floor((al.astype(np.float64) + np.pi) / (2 * np.pi / 30) + 1e-4).astype(int) >= ai.astype(int) - 1).all()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref missing")
def test_model_description_and_voting_scheme_match_reference():
    from objective_slam_b200 import operators as ops, synth
    from oracle import refgpu
    mp, mn = synth.make_model(300, seed=9)
    sp, sn, T = synth.make_scene(mp, mn, 500, seed=10)
    model, d_dist, d_angle = ops.model_description(mp, mn)
    centre = (mp.min(0) + mp.max(0)) / 2
    assert abs(d_dist - 0.1 * np.linalg.norm(mp - centre, axis=1).max()) < 1e-4          # model_description.m:5-13
    assert np.float32(d_angle).view(np.uint32) == 0x3E567750
    res = ops.voting_scheme(model, mp, mn, sp, sn, d_dist, d_angle)                        # skip = 5
    r = refgpu.RefModel(mp, mn, d_dist).lookup(refgpu.RefScene(sp, sn, d_dist, 5))
    assert (r["votes"] == res.votes).all() and (r["counts"] == res.voteCounts).all()
    assert (r["pose"].view(np.uint32) == res.pose.view(np.uint32)).all()
