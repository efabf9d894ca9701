"""Size-independent properties of the CUDA path at sizes the reference cannot run (it needs ~60*N_s^2
bytes and overflows int at N > 46,340): sharding invariance, chunking invariance, determinism,
rigid-transform recovery (the reference's own acceptance test, alignment.cpp:317-323)."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(nm, ns, seed=0):
    from objective_slam_b200 import synth
    mp, mn = synth.make_model(nm, seed=synth.SEED_BASE + seed)
    sp, sn, T = synth.make_scene(mp, mn, ns, seed=synth.SEED_BASE + seed + 1)
    return mp, mn, sp, sn, synth.d_dist_for(mp), T


def _ht_dist(A, B):
    """linalg.cu:9-20: (|dt|, |angle|) between two 4x4 poses."""
    dt = np.linalg.norm(A[:3, 3] - B[:3, 3])
    R = A[:3, :3].T @ B[:3, :3]
    ang = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    return dt, abs(ang)


@pytest.mark.parametrize("nm,ns,df", [(2000, 16000, 5), (10000, 50000, 50)])
def test_rigid_transform_recovery(nm, ns, df):
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, T = _case(nm, ns, seed=nm)
    r = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, df), arrays=False)
    dt, ang = _ht_dist(r.pose.astype(np.float64), T)
    assert dt < 0.1 * 100.0 and ang < np.radians(12), (dt, ang)      # alignment.cpp:141-144 defaults
    assert r.num_scene_pairs == ((ns + df - 1) // df) * ns


def test_sharding_is_invariant():
    """Votes of 4 shards of the reference points add up to the unsharded run; survivors are identical."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import _capi as C
    mp, mn, sp, sn, d, _ = _case(1500, 6000, seed=7)
    m, s = ppf.Model(mp, mn, d), ppf.Scene(sp, sn, d, 3)
    whole = m.ppf_lookup(s)
    lk = ppf.Lookup()
    votes = cells = pairs = 0
    gmax = 0
    for r in range(4):
        C.check(C.lib.ppf_lookup_vote(m._h, s._h, 3, r, 4, lk._h))
        st = lk.stats()
        votes += st.num_nonunique_votes; cells += st.num_unique_votes; pairs += st.num_scene_pairs
        gmax = max(gmax, st.max_vote_count)
    assert (votes, cells, pairs, gmax) == (whole.num_nonunique_votes, whole.num_unique_votes,
                                           whole.num_scene_pairs, whole.max_vote_count)
    surv = {}
    for r in range(4):
        C.check(C.lib.ppf_lookup_vote(m._h, s._h, 3, r, 4, lk._h))
        C.check(C.lib.ppf_lookup_finalize(m._h, gmax, lk._h))
        C.check(C.lib.ppf_lookup_poses(m._h, s._h, lk._h)); C.check(C.lib.ppf_lookup_cluster(m._h, lk._h))
        part = lk.result()
        surv.update(zip(part.votes.tolist(), part.voteCounts.tolist()))
    assert surv == dict(zip(whole.votes.tolist(), whole.voteCounts.tolist()))


def test_chunking_is_invariant_and_runs_are_deterministic(monkeypatch):
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, _ = _case(3000, 5000, seed=3)
    s = ppf.Scene(sp, sn, d, 10)
    a = ppf.Model(mp, mn, d).ppf_lookup(s)
    b = ppf.Model(mp, mn, d).ppf_lookup(s)
    monkeypatch.setenv("PPF_B200_CHUNK_ROWS", "416")
    c = ppf.Model(mp, mn, d).ppf_lookup(s)
    for x in (b, c):
        assert (a.votes == x.votes).all() and (a.voteCounts == x.voteCounts).all()
        assert a.num_nonunique_votes == x.num_nonunique_votes and a.num_unique_votes == x.num_unique_votes
        assert (a.pose.view(np.uint32) == x.pose.view(np.uint32)).all() and a.max_idx == x.max_idx


def test_histogram_checksum_and_model_reuse():
    """sum of all accumulator cells == votes cast; one model handle serves many scenes."""
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, _ = _case(400, 900, seed=11)
    m = ppf.Model(mp, mn, d)
    for df in (1, 2, 9):
        s = ppf.Scene(sp, sn, d, df)
        codes, counts = m.vote_histogram(s)
        r = m.ppf_lookup(s)
        assert int(counts.astype(np.int64).sum()) == r.num_nonunique_votes and len(codes) == r.num_unique_votes
        assert (np.diff(codes.astype(np.int64)) > 0).all()
        assert int(counts.max()) == r.max_vote_count == int(r.voteCounts[0])
        assert ((codes >> np.uint64(32)).astype(np.int64) % df == 0).all()      # only reference rows vote
        assert ((codes & np.uint64(63)) <= 30).all()                               # alpha_idx in [0, 30]


def test_cpu_clustering_variant_agrees_with_gpu_clustering():
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, T = _case(1200, 3000, seed=13)
    s = ppf.Scene(sp, sn, d, 5)
    a = ppf.Model(mp, mn, d).ppf_lookup(s, arrays=False)
    b = ppf.Model(mp, mn, d, cpu_clustering=True).ppf_lookup(s, arrays=False)
    for r in (a, b):
        dt, ang = _ht_dist(r.pose.astype(np.float64), T)
        assert dt < 10.0 and ang < np.radians(12)
    assert abs(np.linalg.det(b.pose[:3, :3].astype(np.float64)) - 1) < 1e-4


def test_config4_multi_model_database():
    """BASELINE configs[3] (reduced): several models with different d_dist against one scene through
    ppf_registration; each pose must equal the object-level lookup of that (scene, model) pair."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    models, dd = [], []
    for k in range(4):
        mp, mn = synth.make_model(600 + 100 * k, seed=900 + k)
        models.append((mp, mn)); dd.append(synth.d_dist_for(mp, 0.05 + 0.01 * k))
    sp, sn, T = synth.make_scene(models[1][0], models[1][1], 20000, seed=77)
    poses, status = ppf.ppf_registration([(sp, sn)], models, dd, ref_point_downsample_factor=10)
    for j, ((mp, mn), d) in enumerate(zip(models, dd)):
        r = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, 10), arrays=False)
        assert status[0, j] == r.status
        assert (poses[0, j].view(np.uint32) == r.pose.view(np.uint32)).all()
    dt, ang = _ht_dist(poses[0, 1].astype(np.float64), T)       # the planted model is found
    assert dt < 10.0 and ang < np.radians(12)


def test_config5_dense_million_point_scene():
    """BASELINE configs[4] (reference points subsampled): a 1M-point TSDF-like lattice cloud.  The reference
    overflows int at 46,341 points; here pair counts are 64-bit and nothing of size N_s^2 exists."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth, _capi as C
    sp, sn = synth.make_lattice_scene(1_000_000, pitch=0.5)
    mp, mn = synth.make_lattice_scene(2000, pitch=0.5, seed=3)
    d = synth.d_dist_for(mp)
    m, s = ppf.Model(mp, mn, d), ppf.Scene(sp, sn, d, 20000)
    whole = m.ppf_lookup(s, arrays=True)
    assert whole.num_scene_pairs == 50 * 1_000_000 and whole.num_nonunique_votes > 0
    lk = ppf.Lookup()
    votes = 0
    for r in range(2):
        C.check(C.lib.ppf_lookup_vote(m._h, s._h, 20000, r, 2, lk._h))
        votes += lk.stats().num_nonunique_votes
    assert votes == whole.num_nonunique_votes
    assert (whole.votes >> np.uint64(32)).max() < 1_000_000


def test_saved_model_is_the_built_model(tmp_path):
    """a model table written by ppf_model_save and read back by ppf_model_load gives the
    same hash arrays, the same layout and bit-identical lookup results as the table that was built."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    mp, mn = synth.make_model(700, seed=31)
    sp, sn, _ = synth.make_scene(mp, mn, 1500, seed=32)
    d = synth.d_dist_for(mp)
    built = ppf.Model(mp, mn, d)
    path = str(tmp_path / "model.ppfb200")
    built.save(path)
    loaded = ppf.Model.load(path)
    assert loaded.n == built.n and loaded.layout() == built.layout()
    for a, b in zip(built.table(), loaded.table()):
        assert (a == b).all()
    s = ppf.Scene(sp, sn, d, 3)
    r0, r1 = built.ppf_lookup(s), loaded.ppf_lookup(s)
    assert r0.num_nonunique_votes == r1.num_nonunique_votes and r0.num_top_votes == r1.num_top_votes
    assert (r0.votes == r1.votes).all() and (r0.voteCounts == r1.voteCounts).all()
    assert (r0.pose.view(np.uint32) == r1.pose.view(np.uint32)).all()
    # a truncated file is refused, not read
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[: len(blob) // 2])
    with pytest.raises(ppf.PpfError):
        ppf.Model.load(path)


def test_config2_both_vote_kernels_agree_at_full_size(monkeypatch):
    """BASELINE configs[1] at full size (10k-point model table, 50k-point scene; every 25th scene point a
    reference point): the grouped kernel and the one-hit-per-pass kernel cast the same 4.8e11 votes into the same
    cells -- vote total, non-zero cells, maximum, every survivor and its count, and the final pose are equal."""
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, T = _case(10000, 50000, seed=2)
    out = {}
    for kind in ("grouped", "classic"):
        monkeypatch.setenv("PPF_B200_VOTE", kind)
        m = ppf.Model(mp, mn, d)
        assert m.layout()[2] == (kind == "grouped")
        out[kind] = m.ppf_lookup(ppf.Scene(sp, sn, d, 25))
        m.close()
    a, b = out["grouped"], out["classic"]
    assert a.num_nonunique_votes == b.num_nonunique_votes > 4e11
    assert (a.num_unique_votes, a.max_vote_count, a.num_top_votes) == (b.num_unique_votes, b.max_vote_count, b.num_top_votes)
    assert (a.votes == b.votes).all() and (a.voteCounts == b.voteCounts).all()
    assert (a.pose.view(np.uint32) == b.pose.view(np.uint32)).all()
    dt, ang = _ht_dist(a.pose.astype(np.float64), T)
    assert dt < 0.1 * 100.0 and ang < np.radians(12), (dt, ang)


def test_scene_size_hint_selects_the_layout_not_the_result():
    """ppf_set_expected_scene_points: a small model meant for big scenes gets the grouped layout; the lookup
    result is the same either way (the vote_kernel fixture is overridden by the hint only when it says auto)."""
    import os
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, _ = _case(500, 3000, seed=11)
    forced = os.environ.pop("PPF_B200_VOTE", None)
    try:
        small = ppf.Model(mp, mn, d)
        big = ppf.Model(mp, mn, d, expected_scene_points=1_000_000)
        assert not small.layout()[2] and big.layout()[2]
        s = ppf.Scene(sp, sn, d, 4)
        a, b = small.ppf_lookup(s), big.ppf_lookup(s)
        assert (a.votes == b.votes).all() and (a.voteCounts == b.voteCounts).all()
        assert (a.pose.view(np.uint32) == b.pose.view(np.uint32)).all()
    finally:
        if forced is not None:
            os.environ["PPF_B200_VOTE"] = forced


def test_sharded_clustering_equals_unsharded():
    """Multi-GPU path: clustering scores computed in 3 interleaved slices and summed are the unsharded scores,
    bit for bit, and the winner is the same (one GPU plays the three ranks in turn)."""
    import ctypes
    import torch
    import objective_slam_b200 as ppf
    from objective_slam_b200 import _capi as C
    mp, mn, sp, sn, d, _ = _case(1200, 5000, seed=17)
    m, s = ppf.Model(mp, mn, d), ppf.Scene(sp, sn, d, 2)
    whole = m.ppf_lookup(s)
    K = whole.num_top_votes
    assert K > 100
    lk = m._lookup if getattr(m, "_lookup", None) is not None else ppf.Lookup()
    C.check(C.lib.ppf_lookup_vote(m._h, s._h, 2, 0, 1, lk._h))
    C.check(C.lib.ppf_lookup_finalize(m._h, whole.max_vote_count, lk._h))
    C.check(C.lib.ppf_lookup_poses(m._h, s._h, lk._h))
    total = torch.zeros(K, dtype=torch.float32, device="cuda")
    part = torch.zeros(K, dtype=torch.float32, device="cuda")
    for r in range(3):
        C.check(C.lib.ppf_lookup_cluster_shard(m._h, lk._h, r, 3))
        C.check(C.lib.ppf_lookup_copy_scores(lk._h, part.data_ptr()))
        assert int((part != 0).sum()) == len(range(r, K, 3))
        total += part
    torch.cuda.synchronize()
    C.check(C.lib.ppf_lookup_set_scores(lk._h, total.data_ptr()))
    C.check(C.lib.ppf_lookup_cluster_finish(lk._h))
    got = lk.result()
    assert got.max_idx == whole.max_idx
    assert (got.vote_counts_out.view(np.uint32) == whole.vote_counts_out.view(np.uint32)).all()
    assert (got.pose.view(np.uint32) == whole.pose.view(np.uint32)).all()
