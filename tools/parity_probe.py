"""Stage-by-stage parity of libppf_b200 against the reference kernels (oracle/_ref) on one small case."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
from oracle import refgpu

def bits(a): return np.ascontiguousarray(a).view(np.uint32)

nm, ns = int(sys.argv[1]) if len(sys.argv) > 1 else 400, int(sys.argv[2]) if len(sys.argv) > 2 else 800
mp, mn = synth.make_model(nm)
sp, sn, T = synth.make_scene(mp, mn, ns)
d = synth.d_dist_for(mp)
ok = True
for df in (1, 5):
    rm = refgpu.RefModel(mp, mn, d); rs = refgpu.RefScene(sp, sn, d, df)
    m = ppf.Model(mp, mn, d); s = ppf.Scene(sp, sn, d, df)
    # features + keys
    rp, rk = rs.features(); p, k = s.features()
    fe = (bits(rp) != bits(p)).any(axis=-1).sum(); ke = (rk != k).sum()
    print(f"df={df} scene features mismatching pairs: {fe}  keys: {ke}  of {rk.size}")
    rmp = rm.features(); pm, km = m.features()
    print(f"      model features mismatching pairs: {(bits(rmp) != bits(pm)).any(axis=-1).sum()}")
    # table
    rt = rm.table(); t = m.table()
    for name, a, b in zip(("hashkeys", "counts", "first", "map"), rt, t):
        same = a.shape == b.shape and (a == b).all()
        ok &= bool(same)
        print(f"      table {name}: {'OK' if same else 'MISMATCH'} {a.shape} {b.shape}")
    # vote histogram
    rc, rn = rm.vote_histogram(rs); c, n = m.vote_histogram(s)
    same = rc.shape == c.shape and (rc == c).all() and (rn == n).all()
    ok &= bool(same)
    print(f"      vote histogram: {'OK' if same else 'MISMATCH'} cells {len(rc)} vs {len(c)} votes {rn.sum()} vs {n.sum()}")
    if not same and len(rc) and len(c):
        da = dict(zip(rc.tolist(), rn.tolist())); db = dict(zip(c.tolist(), n.tolist()))
        diff = [(kk, da.get(kk, 0), db.get(kk, 0)) for kk in set(da) | set(db) if da.get(kk, 0) != db.get(kk, 0)]
        print("      differing cells:", len(diff), diff[:5])
    # lookup
    r = rm.lookup(rs); q = m.ppf_lookup(s)
    print(f"      ref: K={r['K']} votes={r['num_nonunique_votes']} unique={r['num_unique_votes']} max_idx={r['max_idx']}")
    print(f"      b200: K={q.num_top_votes} votes={q.num_nonunique_votes} unique={q.num_unique_votes} max_idx={q.max_idx} exact_alpha={q.num_exact_alpha} ms_vote={q.ms_vote:.3f}")
    same = (r['K'] == q.num_top_votes and (r['votes'] == q.votes).all() and (r['counts'] == q.voteCounts).all())
    ok &= bool(same)
    print(f"      survivors: {'OK' if same else 'MISMATCH'}")
    if same:
        for name, a, b in (("transformations", r['transformations'], q.transformations), ("weighted", r['weighted'], q.weightedVoteCounts),
                           ("trans", r['trans'], q.transformation_trans), ("rots", r['rots'], q.transformation_rots),
                           ("scores", r['scores'], q.vote_counts_out), ("pose", r['pose'], q.pose)):
            nb = (bits(a) != bits(b)).sum()
            print(f"      {name}: bit mismatches {nb} of {a.size}  maxabs {np.abs(a - b).max() if a.size else 0}")
            ok &= nb == 0
        ok &= r['max_idx'] == q.max_idx
print("PARITY", "OK" if ok else "FAILED")
