"""Small driver for ncu: one model build + a few lookups.  usage: profile_vote.py n_model n_scene ref_df [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
nm, ns, df = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
SEED = 0xD205 + 2
mp, mn = synth.make_model(nm, seed=SEED)
sp, sn, T = synth.make_scene(mp, mn, ns, seed=SEED + 1)
d = synth.d_dist_for(mp)
m = ppf.Model(mp, mn, d)
hk, cnt, first, mapp = m.table()
print("U", len(hk), "max bucket", int(cnt.max()), "mean bucket", float(cnt.mean()))
s = ppf.Scene(sp, sn, d, df)
for i in range(reps):
    r = m.ppf_lookup(s, arrays=False)
    print(f"pairs {r.num_scene_pairs} votes {r.num_nonunique_votes} ms_vote {r.ms_vote:.3f} votes/s {r.num_nonunique_votes / r.ms_vote * 1e3:.3e} "
          f"K {r.num_top_votes} exact {r.num_exact_alpha} fin {r.ms_finalize:.2f} pc {r.ms_pose_cluster:.2f}")
