"""CUDA product against the plain-C CPU oracle (oracle/ppf_oracle.c).  The CPU uses libm, the GPU its
approximate sqrt/div/acos, so features may flip a bin for pairs on an edge: the rates are bounded,
and every integer stage is exact once both sides use the same keys."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cuda_matches_cpu_oracle():
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    from oracle import cpu
    mp, mn = synth.make_model(300, seed=77)
    sp, sn, T = synth.make_scene(mp, mn, 500, seed=78)
    d = synth.d_dist_for(mp)
    m, s = ppf.Model(mp, mn, d), ppf.Scene(sp, sn, d, 2)
    _, gk = s.features()
    _, ck = cpu.scene_features(sp, sn, d, 2)
    assert ((gk == 0) == (ck == 0)).all()
    assert (gk != ck).mean() < 1e-3
    _, mk = m.features()
    hk, cnt, first, mapp = m.table()
    ohk, ocnt, ofirst, omap = cpu.hash_array(mk)                      # same keys -> identical table
    assert (hk == ohk).all() and (cnt == ocnt).all() and (first == ofirst).all() and (mapp == omap).all()
    o = cpu.lookup(mp, mn, sp, sn, d, 2, model_keys=mk, scene_keys=gk, histogram=True)
    q = m.ppf_lookup(s)
    assert o["num_nonunique_votes"] == q.num_nonunique_votes           # probe + bucket sizes exact
    codes, counts = m.vote_histogram(s)
    a = dict(zip(codes.tolist(), counts.tolist())); b = dict(zip(o["hist_codes"].tolist(), o["hist_counts"].tolist()))
    l1 = sum(abs(a.get(k, 0) - b.get(k, 0)) for k in set(a) | set(b))
    assert l1 <= 2e-3 * q.num_nonunique_votes
    # poses of the common survivors within the north-star tolerance
    common = sorted(set(q.votes.tolist()) & set(o["votes"].tolist()))
    assert len(common) >= 0.9 * q.num_top_votes
    qi = {v: i for i, v in enumerate(q.votes.tolist())}; oi = {v: i for i, v in enumerate(o["votes"].tolist())}
    for v in common:
        A, B = q.transformations[qi[v]], o["transformations"][oi[v]]
        assert np.linalg.norm(A[:3, :3] - B[:3, :3]) < 1e-4 and np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-4 * 100
