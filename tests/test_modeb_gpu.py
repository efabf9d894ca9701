"""Full-size parity through oracle Mode B.

The whole-kernel replay of the reference (Mode A, oracle/ref_harness.cu) needs ~60 N_s^2 bytes + 8 B per vote and
overflows `int` beyond 46,340 points, so it stops near 20k-point scenes.  Mode B streams the scene for a SAMPLE of
reference points through the reference's own device functions (compute_ppf, disc_feature, hash,
trans_model_scene, linked with -rdc) against the reference's own ParallelHashArray, into one dense (m_r, alpha)
histogram per reference point.  Reference points are independent (the high 32 bits of a vote code are s_r), so
equality of every accumulator cell of the sampled reference points pins the vote stage of the product at the
BASELINE.json sizes: configs[1] (10k-point model, 50k-point scene), configs[3] (multi-model database, 200k-point
scene) and configs[4] (1M-point dense scene).  First: Mode B == Mode A, cell for cell, wherever both run."""
import numpy as np
import pytest

from conftest import have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref missing")]


def _sample(ns, df, n_refs):
    """(shard_rank, shard_count, refs): ~n_refs reference points spread over the scene, as one shard."""
    R_all = (ns + df - 1) // df
    count = max(1, R_all // n_refs)
    rank = count // 2
    return rank, count, np.arange(rank, R_all, count, dtype=np.int64) * df


def _check_against_mode_b(mp, mn, sp, sn, d, df, n_refs, expected_scene_points=0):
    import objective_slam_b200 as ppf
    from oracle import refgpu
    rank, count, refs = _sample(len(sp), df, n_refs)
    rm = refgpu.RefModel(mp, mn, d)
    rc, rn, votes = rm.modeb_histogram(sp, sn, refs)
    rm.close()
    m = ppf.Model(mp, mn, d, expected_scene_points=expected_scene_points)
    s = ppf.Scene(sp, sn, d, df)
    c, n = m.vote_histogram(s, rank, count)
    assert votes > 0 and len(rc) > 0
    assert int(n.astype(np.uint64).sum()) == votes, "votes cast for the sampled reference points"
    assert rc.shape == c.shape and (rc == c).all() and (rn == n).all(), "every accumulator cell of the sample"
    return m, s, votes, len(refs)


@pytest.mark.parametrize("nm,ns,df,kind", [(640, 1500, 1, "lumpy"), (300, 500, 5, "lumpy"), (200, 420, 1, "lattice"),
                                           (200, 420, 2, "degenerate")])
def test_mode_b_equals_mode_a(nm, ns, df, kind):
    """The tiled oracle reproduces the whole-kernel replay (the reference's __global__ kernels unmodified): the
    external-linkage copies of the device functions give the same bits as the copies inlined into the kernels."""
    from test_parity_gpu import clouds
    from oracle import refgpu
    mp, mn, sp, sn, d, _ = clouds(nm, ns, 0.05 if kind == "lumpy" else 0.08, seed=nm + ns, kind=kind)
    rm = refgpu.RefModel(mp, mn, d); rs = refgpu.RefScene(sp, sn, d, df)
    ac, an = rm.vote_histogram(rs)
    refs = np.arange(0, len(sp), df)
    pick = refs[:: max(1, len(refs) // 40)]
    bc, bn, votes = rm.modeb_histogram(sp, sn, pick)
    keep = np.isin((ac >> np.uint64(32)).astype(np.int64), pick)
    assert keep.any()
    assert (ac[keep] == bc).all() and (an[keep] == bn).all()
    assert votes == int(an[keep].astype(np.uint64).sum())
    # all reference points at once = the complete Mode A histogram
    bc, bn, _ = rm.modeb_histogram(sp, sn, refs)
    assert bc.shape == ac.shape and (bc == ac).all() and (bn == an).all()


def test_config1_full_size_10k_model_50k_scene():
    """BASELINE configs[1], the benchmark workload itself (bench.py: seed 0xD205+2, tau_d 0.05, ref_point_df 8):
    every accumulator cell of 32 reference points spread over the scene, on the real 10k-row / 16-chunk table."""
    from objective_slam_b200 import synth
    seed = 0xD205 + 2
    mp, mn = synth.make_model(10000, seed=seed)
    sp, sn, _ = synth.make_scene(mp, mn, 50000, seed=seed + 1)
    m, s, votes, n = _check_against_mode_b(mp, mn, sp, sn, synth.d_dist_for(mp, 0.05), 8, 32)
    assert m.layout()[0] >= 7 and votes > 1e9 and n >= 32


def test_config4_dense_1m_point_scene():
    """BASELINE configs[4]: 2k-point model, 1M-point dense scene (room lattice + one object, tools/profile_big_scene.py):
    8 reference points; the reference itself cannot index this scene (int overflow at N > 46,340)."""
    from objective_slam_b200 import synth
    ns = 1_000_000
    mp, mn = synth.make_model(2000, seed=0xD209)
    sp0, sn0, _ = synth.make_scene(mp, mn, 20000, seed=0xD20A)
    lp, ln = synth.make_lattice_scene(ns - 20000, pitch=1.0)
    sp = np.concatenate([sp0, lp + sp0.min(0)]).astype(np.float32)
    sn = np.concatenate([sn0, ln]).astype(np.float32)
    perm = np.random.default_rng(1).permutation(len(sp))
    sp, sn = sp[perm], sn[perm]
    _check_against_mode_b(mp, mn, sp, sn, synth.d_dist_for(mp), 100, 8, expected_scene_points=ns)


@pytest.mark.parametrize("j", [0, 7, 19])
def test_config3_multi_model_database_200k_scene(j):
    """BASELINE configs[3]: 20 models of different size / d_dist against one 200k-point scene; models 0, 7 and 19,
    6 reference points each."""
    from objective_slam_b200 import synth
    models = [synth.make_model(2000, seed=0xD300 + k, diameter=60.0 + 4.0 * k) for k in range(20)]
    rng = np.random.default_rng(0xD3FF)
    parts = []
    for k in (0, 7, 19):                                              # three of the objects are in the scene
        p, n, _ = synth.make_scene(models[k][0], models[k][1], 3000, seed=0xD340 + k)
        parts.append((p + rng.random(3) * 200.0, n))
    lp, ln = synth.make_lattice_scene(200_000 - 9000, pitch=1.5)
    sp = np.concatenate([p for p, _ in parts] + [lp]).astype(np.float32)
    sn = np.concatenate([n for _, n in parts] + [ln]).astype(np.float32)
    perm = rng.permutation(len(sp))
    sp, sn = sp[perm], sn[perm]
    mp, mn = models[j]
    _check_against_mode_b(mp, mn, sp, sn, synth.d_dist_for(mp, 0.05), 40, 6, expected_scene_points=len(sp))
