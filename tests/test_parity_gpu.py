"""Parity proper: libppf_b200 (through the C ABI) against the reference's own CUDA kernels
(oracle/_ref: kernel.cu + parallel_hash_array.hpp compiled for sm_100a, replayed by
oracle/ref_harness.cu) on the same seeded inputs.  Everything is compared BIT-EXACTLY: quantised
features, keys, hash-table arrays, every accumulator cell, survivors and their order, poses,
quaternions, cluster scores, winner.  (The north star only asks 1e-4 for poses; we get 0.)"""
import numpy as np
import pytest

from conftest import golden, have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref missing")]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def clouds(nm, ns, tau=0.05, seed=0, kind="lumpy"):
    from objective_slam_b200 import synth
    mp, mn = synth.make_model(nm, seed=100 + seed)
    if kind == "lumpy":
        sp, sn, T = synth.make_scene(mp, mn, ns, seed=200 + seed)
    elif kind == "lattice":       # TSDF-like lattice, axis-aligned normals: exact zeros everywhere
        sp, sn = synth.make_lattice_scene(ns, pitch=2.5, seed=seed)
        mp, mn = synth.make_lattice_scene(nm, pitch=2.5, seed=seed + 1)
        T = np.eye(4)
    elif kind == "degenerate":    # duplicate points, zero / parallel normals, coincident pairs
        sp, sn, T = synth.make_scene(mp, mn, ns, seed=200 + seed)
        sp[5] = sp[6]; sn[7] = 0; sn[8] = sn[9]; sp[10] = sp[11]; sn[10] = sn[11]
        mp[3] = mp[4]; mn[5] = 0; mn[6] = mn[7] = np.array([0, 0, 1], np.float32)
        sn[12] = np.array([1, 0, 0], np.float32); sn[13] = np.array([0, 1, 0], np.float32)
        sn[14] = np.array([0, -1, 0], np.float32); sn[15] = np.array([-1, 0, 0], np.float32)
    return mp, mn, sp, sn, synth.d_dist_for(mp, tau), T


def compare_all(mp, mn, sp, sn, d, df, thr=0.4, l1=False, avg=False, check_hist=True):
    import objective_slam_b200 as ppf
    from oracle import refgpu
    rm = refgpu.RefModel(mp, mn, d); rs = refgpu.RefScene(sp, sn, d, df)
    m = ppf.Model(mp, mn, d, thr, use_l1_norm=l1, use_averaged_clusters=avg); s = ppf.Scene(sp, sn, d, df)
    # point_pair_feature / my_discretize / hash
    rp, rk = rs.features(); p, k = s.features()
    assert (bits(rp) == bits(p)).all(), "scene quantised features"
    assert (rk == k).all(), "scene keys"
    pm, km = m.features()
    assert (bits(rm.features()) == bits(pm)).all(), "model quantised features"
    # model_description
    for name, a, b in zip(("hashkeys", "counts", "first", "map"), rm.table(), m.table()):
        assert a.shape == b.shape and (a == b).all(), f"table {name}"
    # voting_scheme: every accumulator cell
    if check_hist:
        rc, rn = rm.vote_histogram(rs); c, n = m.vote_histogram(s)
        assert rc.shape == c.shape and (rc == c).all() and (rn == n).all(), "vote histogram"
    r = rm.lookup(rs, thr, l1, avg); q = m.ppf_lookup(s)
    assert r["num_nonunique_votes"] == q.num_nonunique_votes
    assert r["num_unique_votes"] == q.num_unique_votes
    if r["K"] < 0:
        assert q.num_top_votes == 0 and q.status != 0
        return r, q
    assert r["K"] == q.num_top_votes
    assert (r["votes"] == q.votes).all() and (r["counts"] == q.voteCounts).all(), "survivors / order"
    assert (bits(r["transformations"]) == bits(q.transformations)).all(), "poses"
    assert (bits(r["weighted"]) == bits(q.weightedVoteCounts)).all()
    assert (bits(r["rots"]) == bits(q.transformation_rots)).all(), "quaternions"
    if not avg:       # the reference's in-place averaging races (kernel.cu:758 vs :742,750)
        assert (bits(r["trans"]) == bits(q.transformation_trans)).all()
        assert (bits(r["scores"]) == bits(q.vote_counts_out)).all(), "cluster scores"
        assert r["max_idx"] == q.max_idx
        assert (bits(r["pose"]) == bits(q.pose)).all(), "final pose"
    return r, q


@pytest.mark.parametrize("nm,ns,df,tau", [(300, 500, 1, 0.05), (300, 500, 5, 0.05), (257, 333, 3, 0.1),
                                          (640, 1500, 1, 0.05), (1000, 1000, 5, 0.05), (33, 2100, 7, 0.05)])
def test_stagewise_parity(nm, ns, df, tau):
    mp, mn, sp, sn, d, _ = clouds(nm, ns, tau, seed=nm + ns)
    compare_all(mp, mn, sp, sn, d, df)


@pytest.mark.parametrize("kind", ["lattice", "degenerate"])
@pytest.mark.parametrize("df", [1, 4])
def test_degenerate_inputs(kind, df):
    """NaN features (acos > 1, 0/0), coincident points, zero normals, exact zeros / signed zeros."""
    mp, mn, sp, sn, d, _ = clouds(200, 420, 0.08, seed=5, kind=kind)
    r, q = compare_all(mp, mn, sp, sn, d, df)
    assert q.num_nonunique_votes > 0


def test_model_split_into_several_chunks(monkeypatch):
    """2000-point model -> two accumulator chunks; forced 64-row chunks -> 8 chunks.  Same answers."""
    mp, mn, sp, sn, d, _ = clouds(2000, 700, 0.05, seed=9)
    compare_all(mp, mn, sp, sn, d, 5, check_hist=False)
    monkeypatch.setenv("PPF_B200_CHUNK_ROWS", "64")
    mp, mn, sp, sn, d, _ = clouds(500, 600, 0.05, seed=10)
    compare_all(mp, mn, sp, sn, d, 2)


def test_hit_queue_overflow(monkeypatch):
    """Grouped kernel: a reference point with more hits than the shared-memory queue holds re-collects its
    hits chunk by chunk (queue forced down to 4096 records; 6000-point scene, 3 model chunks)."""
    monkeypatch.setenv("PPF_B200_VOTE_QUEUE", "4096")
    monkeypatch.setenv("PPF_B200_CHUNK_ROWS", "256")
    mp, mn, sp, sn, d, _ = clouds(600, 6000, 0.08, seed=12)
    compare_all(mp, mn, sp, sn, d, 40)


def test_candidate_buffer_overflow(monkeypatch):
    """More cells above the running threshold than the candidate buffer holds (forced down to 64): the buffer grows
    to the counted size and ONE more pass, seeded with the first pass's maximum, emits exactly the survivors."""
    monkeypatch.setenv("PPF_B200_CAND_CAP", "64")
    mp, mn, sp, sn, d, _ = clouds(300, 500, 0.05, seed=77)
    compare_all(mp, mn, sp, sn, d, 1, thr=0.1)


@pytest.mark.parametrize("l1,avg,thr", [(True, False, 0.4), (False, True, 0.4), (False, False, 0.9), (False, False, 0.0)])
def test_lookup_options(l1, avg, thr):
    mp, mn, sp, sn, d, _ = clouds(250, 400, 0.06, seed=21)
    compare_all(mp, mn, sp, sn, d, 2, thr=thr, l1=l1, avg=avg, check_hist=False)


@pytest.mark.parametrize("nm,ns", [(1, 50), (2, 50), (50, 1), (50, 2), (3, 3)])
def test_tiny_and_ragged_sizes(nm, ns):
    """count <= 1 makes every reference kernel a no-op (kernel.cu:406,461,...); K <= 1 gives a zero pose."""
    mp, mn, sp, sn, d, _ = clouds(60, 60, 0.1, seed=3)
    compare_all(mp[:nm], mn[:nm], sp[:ns], sn[:ns], d, 1)


def test_scene_far_from_model_has_no_votes():
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, _ = clouds(100, 100, 0.05, seed=4)
    far = (sp * 1000).astype(np.float32)                     # every scene pair is longer than any model pair
    r = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(far, sn, d, 1))
    assert r.status == 4 and r.num_nonunique_votes == 0 and r.num_top_votes == 0 and not r.pose.any()


@pytest.mark.parametrize("name", ["tiny_df1", "tiny_df3"])
def test_golden(name):
    """Against the committed fixtures (reference kernels on a B200, tests/golden/make_golden.py)."""
    import objective_slam_b200 as ppf
    g = golden(name)
    d, df = float(g["d_dist"]), int(g["ref_df"])
    m = ppf.Model(g["model_pts"], g["model_nrm"], d); s = ppf.Scene(g["scene_pts"], g["scene_nrm"], d, df)
    p, k = s.features()
    assert (bits(p) == bits(g["scene_ppf"])).all() and (k == g["scene_keys"]).all()
    hk, cnt, first, mapp = m.table()
    assert (hk == g["hashkeys"]).all() and (cnt == g["counts"]).all() and (mapp == g["map"]).all()
    c, n = m.vote_histogram(s)
    assert (c == g["hist_codes"]).all() and (n == g["hist_counts"]).all()
    q = m.ppf_lookup(s)
    assert (q.votes == g["votes"]).all() and (q.voteCounts == g["vote_counts"]).all()
    assert (bits(q.transformations) == bits(g["transformations"])).all()
    assert (bits(q.vote_counts_out) == bits(g["scores"])).all() and q.max_idx == int(g["max_idx"])
    assert (bits(q.pose) == bits(g["pose"])).all()


def test_registration_boundary():
    """ppf_registration (ppf.h:9-15): 2 scenes x 2 models with different d_dist, host clouds in, poses out."""
    import objective_slam_b200 as ppf
    from oracle import refgpu
    mpa, mna, spa, sna, da, _ = clouds(220, 380, 0.05, seed=31)
    mpb, mnb, spb, snb, db, _ = clouds(180, 300, 0.07, seed=32)
    poses, status = ppf.ppf_registration([(spa, sna), (spb, snb)], [(mpa, mna), (mpb, mnb)], [da, db],
                                         ref_point_downsample_factor=3, devUse=1)
    for i, (sp, sn) in enumerate([(spa, sna), (spb, snb)]):
        for j, (mp, mn, d) in enumerate([(mpa, mna, da), (mpb, mnb, db)]):
            r = refgpu.RefModel(mp, mn, d).lookup(refgpu.RefScene(sp, sn, d, 3))
            if r["K"] > 0:
                assert status[i, j] == 0 and (bits(r["pose"]) == bits(poses[i, j])).all()
            else:
                assert not poses[i, j].any()


@pytest.mark.parametrize("mem", ["host", "device"])
def test_pointnormal_strided_clouds(mem):
    """The layout INTEGRATION.md tells the maintainer to pass: pcl::PointNormal records of 12 floats
    (x y z pad | normal_x normal_y normal_z pad | curvature pad pad pad), xyz_stride = nrm_stride = 12, the normal
    pointer 4 floats into the record -- through the C ABI (ppf_scene_create / ppf_model_create / ppf_registration),
    host and device memory.  Same bits as the packed N x 3 arrays; the padding holds NaNs that must never be read."""
    import ctypes
    import torch
    import objective_slam_b200 as ppf
    from objective_slam_b200 import _capi as C
    mp, mn, sp, sn, d, _ = clouds(230, 410, 0.05, seed=91)

    def records(p, n):
        r = np.full((len(p), 12), np.nan, np.float32)
        r[:, 0:3] = p; r[:, 4:7] = n
        return r

    want = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, 2))
    rm, rs = records(mp, mn), records(sp, sn)
    if mem == "device":
        tm, ts = torch.from_numpy(rm).cuda(), torch.from_numpy(rs).cuda()
        torch.cuda.synchronize()
        pm, ps, kind = tm.data_ptr(), ts.data_ptr(), C.PPF_MEM_DEVICE
    else:
        pm, ps, kind = rm.ctypes.data, rs.ctypes.data, C.PPF_MEM_HOST
    hm, hs, lk = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    C.check(C.lib.ppf_model_create(pm, 12, pm + 16, 12, len(mp), kind, float(d), 0.4, 0, 0, ctypes.byref(hm)))
    C.check(C.lib.ppf_scene_create(ps, 12, ps + 16, 12, len(sp), kind, ctypes.byref(hs)))
    C.check(C.lib.ppf_lookup_create(ctypes.byref(lk)))
    C.check(C.lib.ppf_model_lookup(hm, hs, 2, lk))
    pose = np.zeros((4, 4), np.float32)
    C.check(C.lib.ppf_lookup_get(lk, None, None, None, None, None, None, None, pose.ctypes.data))
    st = C.LookupStats()
    C.check(C.lib.ppf_lookup_get_stats(lk, ctypes.byref(st)))
    assert st.num_nonunique_votes == want.num_nonunique_votes and st.num_top_votes == want.num_top_votes
    assert (bits(pose) == bits(want.pose)).all()
    C.lib.ppf_lookup_destroy(lk); C.lib.ppf_scene_destroy(hs); C.lib.ppf_model_destroy(hm)
    if mem == "host":       # the drop-in call with ppf_cloud_t descriptors of PointNormal records (host clouds only)
        sd = (C.CloudDesc * 1)(C.CloudDesc(ps, 12, ps + 16, 12, len(sp)))
        md = (C.CloudDesc * 1)(C.CloudDesc(pm, 12, pm + 16, 12, len(mp)))
        dd = np.array([d], np.float32)
        out = np.zeros((1, 1, 4, 4), np.float32)
        status = np.zeros(1, np.int32)
        C.check(C.lib.ppf_registration(sd, 1, md, 1, dd.ctypes.data, 2, 0.4, 0, 0, 0, 0, None, out.ctypes.data,
                                       status.ctypes.data))
        assert status[0] == 0 and (bits(out[0, 0]) == bits(want.pose)).all()


def test_device_resident_clouds_match_host_clouds():
    import torch
    import objective_slam_b200 as ppf
    mp, mn, sp, sn, d, _ = clouds(200, 300, 0.05, seed=41)
    a = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, 2))
    t = lambda x: torch.from_numpy(x).cuda()
    b = ppf.Model(t(mp), t(mn), d).ppf_lookup(ppf.Scene(t(sp), t(sn), d, 2))
    assert (a.votes == b.votes).all() and (a.voteCounts == b.voteCounts).all() and (bits(a.pose) == bits(b.pose)).all()


def test_config3_model_build_stress_5k():
    """BASELINE configs[2]: 5k-point model = 25 M pairs, sort-based table vs the reference's ParallelHashArray."""
    import objective_slam_b200 as ppf
    from oracle import refgpu
    mp, mn, _, _, d, _ = clouds(5000, 10, 0.05, seed=55)
    rt = refgpu.RefModel(mp, mn, d).table()
    t = ppf.Model(mp, mn, d).table()
    for name, a, b in zip(("hashkeys", "counts", "first", "map"), rt, t):
        assert a.shape == b.shape and (a == b).all(), name
    assert len(t[3]) == 25_000_000 and t[1][0] == 5000          # bucket 0 = the self pairs


@pytest.mark.parametrize("sort", ["own", "cub"])
@pytest.mark.parametrize("nm,tau", [(1200, 0.004), (700, 0.02), (90, 0.3)])
def test_model_table_radix_passes(monkeypatch, vote_kernel, sort, nm, tau):
    """model_description with few / many buckets: the table is sorted by bucket rank over ceil(log2 U) bits, i.e. one,
    two or three passes of the repo's radix sort (tau_d = 0.004: ~1e5 buckets, 17-18 rank bits, odd pass count, the
    payload ping-pong ends in the map either way), and the same with the library sort behind the A/B hook.
    hashkeys / counts / firstHashkeyIndex / hashkeyToDataMap equal the reference's ParallelHashArray."""
    if vote_kernel != "grouped":
        pytest.skip("the table build does not depend on the vote kernel")
    import objective_slam_b200 as ppf
    from oracle import refgpu
    monkeypatch.setenv("PPF_B200_SORT", sort)
    mp, mn, _, _, d, _ = clouds(nm, 10, tau, seed=61 + nm)
    rt = refgpu.RefModel(mp, mn, d).table()
    t = ppf.Model(mp, mn, d).table()
    for name, a, b in zip(("hashkeys", "counts", "first", "map"), rt, t):
        assert a.shape == b.shape and (a == b).all(), name
    if tau == 0.004:
        assert len(t[0]) > 65536                                     # three radix passes


@pytest.mark.parametrize("tau", [0.0849, 0.0814, 0.0881])
def test_far_cell_collision(tau):
    """kernel.cu:460-501: ANY scene key equal to a model key votes -- also the key of a feature cell whose distance
    bin lies beyond every model pair (kd >= K_d), when its 32-bit FNV collides with a model key.  For
    make_model(300, seed=105) and these tau_d such cells exist and are geometrically reachable
    (tools/find_far_collision.py): plant a scene pair in each and expect the reference's extra votes."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import operators, synth
    from oracle import farcells, refgpu
    mp, mn = synth.make_model(300, seed=105)
    d = synth.d_dist_for(mp, tau)
    sp, sn, _ = synth.make_scene(mp, mn, 500, seed=205)
    m = ppf.Model(mp, mn, d)
    hk, cnt, _, _ = m.table()
    pm, _ = m.features()
    K_d = int(np.nanmax(pm[..., 0]) / np.float32(d) + 0.5) + 1                 # one past the model's last distance bin
    planted, expect = [], 0
    for cell, key in farcells.far_collisions(hk, K_d, 4 * K_d, d):
        pair = farcells.plant_pair(cell, d, origin=(3.0, 3.0 + 9.0 * len(planted), 3.0))
        if pair is None:
            continue
        p1, n1, p2, n2 = [np.asarray(x, np.float32) for x in pair]
        _, k = operators.discretized_pair_feature(p1, n1, p2, n2, d)
        assert int(k[0]) == key, "the planted pair must land in the colliding far cell"
        planted.append((p1, n1, p2, n2))
        expect += int(cnt[np.searchsorted(hk, np.uint32(key))])
    assert planted, "no reachable far-cell collision for this model / tau_d (see tools/find_far_collision.py)"
    pp = np.concatenate([[p[0], p[2]] for p in planted]).astype(np.float32)
    pn = np.concatenate([[p[1], p[3]] for p in planted]).astype(np.float32)
    # the planted pairs alone: every vote comes through a far cell (the points are > K_d bins apart)
    if len(pp) == 2:
        r = refgpu.RefModel(mp, mn, d); s2 = refgpu.RefScene(pp, pn, d, 1)
        assert r.lookup(s2)["num_nonunique_votes"] >= expect > 0
        q = m.ppf_lookup(ppf.Scene(pp, pn, d, 1))
        assert q.num_nonunique_votes == r.lookup(s2)["num_nonunique_votes"]
        compare_all(mp, mn, pp, pn, d, 1)
    # inside a cluttered scene: every accumulator cell, survivor, pose equal to the reference's
    sp2, sn2 = np.concatenate([sp, pp]), np.concatenate([sn, pn])
    r, q = compare_all(mp, mn, sp2, sn2, d, 1)
    base = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, 1))
    assert q.num_nonunique_votes >= base.num_nonunique_votes + expect
