"""The multi-GPU path behind the C ABI (ppf_model_lookup_sharded / ppf_registration_sharded, include/ppf_b200.h) on
ONE GPU: `world` host threads, one rank each, coupled by ppf_comm_create_local (NCCL refuses two ranks on one
device; bench.py --gpus N runs the same entry points over NCCL and reports "parity_sharded").  Every rank must return
what the single-GPU lookup returns, bit for bit: survivors and their order, poses, cluster scores, winner."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _run_ranks(world, fn):
    """fn(rank, comm) in `world` threads; returns the per-rank results, re-raises the first exception."""
    import torch
    from objective_slam_b200.dist import Comm
    comms = Comm.local(world)
    out, err = [None] * world, [None] * world
    dev = torch.cuda.current_device()

    def body(r):
        try:
            torch.cuda.set_device(dev)
            out[r] = fn(r, comms[r])
        except BaseException as e:      # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert not any(t.is_alive() for t in th), "a rank hangs in a collective"
    for e in err:
        if e is not None:
            raise e
    return out


def _same(a, b):
    assert a.num_top_votes == b.num_top_votes and a.max_vote_count == b.max_vote_count
    assert (a.votes == b.votes).all() and (a.voteCounts == b.voteCounts).all(), "survivors / order"
    assert (bits(a.transformations) == bits(b.transformations)).all(), "poses"
    assert (bits(a.vote_counts_out) == bits(b.vote_counts_out)).all(), "cluster scores"
    assert a.max_idx == b.max_idx and (bits(a.pose) == bits(b.pose)).all(), "winner"


@pytest.mark.parametrize("world,nm,ns,df,avg", [(2, 600, 2500, 3, False), (3, 300, 1100, 1, False), (4, 1500, 6000, 7, False),
                                                (2, 300, 900, 2, True)])
def test_sharded_lookup_equals_single_gpu(world, nm, ns, df, avg):
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    from objective_slam_b200.dist import lookup_sharded
    mp, mn = synth.make_model(nm, seed=900 + nm)
    sp, sn, _ = synth.make_scene(mp, mn, ns, seed=901 + ns)
    d = synth.d_dist_for(mp)
    whole = ppf.Model(mp, mn, d, use_averaged_clusters=avg).ppf_lookup(ppf.Scene(sp, sn, d, df))
    assert whole.num_top_votes > 1

    def rank_fn(r, comm):
        m = ppf.Model(mp, mn, d, use_averaged_clusters=avg)          # replicated table, one handle per rank
        return lookup_sharded(m, ppf.Scene(sp, sn, d, df), ppf.Lookup(), comm, arrays=True)

    parts = _run_ranks(world, rank_fn)
    for p in parts:
        _same(p, whole)
    assert sum(p.num_nonunique_votes for p in parts) == whole.num_nonunique_votes      # the ranks' own shards add up
    assert sum(p.num_scene_pairs for p in parts) == whole.num_scene_pairs


def test_sharded_lookup_with_a_rank_that_has_no_votes():
    """More ranks than reference points: the empty ranks still take part in every collective."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    from objective_slam_b200.dist import lookup_sharded
    mp, mn = synth.make_model(200, seed=77)
    sp, sn, _ = synth.make_scene(mp, mn, 300, seed=78)
    d = synth.d_dist_for(mp)
    whole = ppf.Model(mp, mn, d).ppf_lookup(ppf.Scene(sp, sn, d, 150))              # 2 reference points
    parts = _run_ranks(4, lambda r, c: lookup_sharded(ppf.Model(mp, mn, d), ppf.Scene(sp, sn, d, 150), ppf.Lookup(), c, True))
    for p in parts:
        _same(p, whole)
    assert [p.num_scene_pairs for p in parts] == [300, 300, 0, 0]


def test_loaded_model_works_sharded(tmp_path):
    """A model read back from the persistent database carries its options (ppf_model_params) and serves the sharded
    path (round-1 advisor finding: Model.load left them unset)."""
    import objective_slam_b200 as ppf
    from objective_slam_b200 import synth
    from objective_slam_b200.dist import lookup_sharded
    mp, mn = synth.make_model(400, seed=81)
    sp, sn, _ = synth.make_scene(mp, mn, 1500, seed=82)
    d = synth.d_dist_for(mp)
    m = ppf.Model(mp, mn, d, vote_count_threshold=0.3)
    path = str(tmp_path / "m.ppf")
    m.save(path)
    whole = m.ppf_lookup(ppf.Scene(sp, sn, d, 2))
    ld = ppf.Model.load(path)
    assert (ld.d_dist, ld.vote_count_threshold, ld.use_averaged_clusters, ld.use_l1_norm) == (m.d_dist, pytest.approx(0.3), False, False)
    parts = _run_ranks(2, lambda r, c: lookup_sharded(ppf.Model.load(path), ppf.Scene(sp, sn, d, 2), ppf.Lookup(), c, True))
    for p in parts:
        _same(p, whole)


def test_registration_sharded_equals_registration():
    """ppf_registration_sharded (every rank, same arguments) == ppf_registration (ppf.h:9-15)."""
    import ctypes
    import objective_slam_b200 as ppf
    from objective_slam_b200 import _capi as C, synth
    mpa, mna = synth.make_model(260, seed=91)
    mpb, mnb = synth.make_model(200, seed=92)
    spa, sna, _ = synth.make_scene(mpa, mna, 700, seed=93)
    da, db = synth.d_dist_for(mpa, 0.05), synth.d_dist_for(mpb, 0.07)
    want, wst = ppf.ppf_registration([(spa, sna)], [(mpa, mna), (mpb, mnb)], [da, db], 2)

    def rank_fn(r, comm):
        keep = [np.ascontiguousarray(x, np.float32) for x in (spa, sna, mpa, mna, mpb, mnb)]
        sd = (C.CloudDesc * 1)(C.CloudDesc(keep[0].ctypes.data, 3, keep[1].ctypes.data, 3, len(spa)))
        md = (C.CloudDesc * 2)(C.CloudDesc(keep[2].ctypes.data, 3, keep[3].ctypes.data, 3, len(mpa)),
                               C.CloudDesc(keep[4].ctypes.data, 3, keep[5].ctypes.data, 3, len(mpb)))
        dd = np.array([da, db], np.float32)
        poses = np.zeros((1, 2, 4, 4), np.float32)
        st = np.zeros((1, 2), np.int32)
        C.check(C.lib.ppf_registration_sharded(sd, 1, md, 2, dd.ctypes.data, 2, 0.4, 0, 0, 0, comm._h, poses.ctypes.data,
                                               st.ctypes.data))
        return poses, st

    for poses, st in _run_ranks(3, rank_fn):
        assert (st == wst).all() and (bits(poses) == bits(want)).all()
