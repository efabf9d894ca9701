"""configs[4] of the bench alone: 2k-point model against the 1M-point scene (room lattice + one object), df 8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import objective_slam_b200 as ppf
synth = bench.load_synth()
mp, mn = synth.make_model(2000, seed=0xD209)
sp0, sn0, T = synth.make_scene(mp, mn, 20000, seed=0xD20A)
lp, ln = synth.make_lattice_scene(1000000 - 20000, pitch=1.0)
sp = np.concatenate([sp0, lp + sp0.min(0)]).astype(np.float32); sn = np.concatenate([sn0, ln]).astype(np.float32)
perm = np.random.default_rng(1).permutation(len(sp)); sp, sn = sp[perm], sn[perm]
d = synth.d_dist_for(mp, 0.05)
m, s = ppf.Model(mp, mn, d, expected_scene_points=len(sp)), ppf.Scene(sp, sn, d, 8)
for i in range(3):
    q = m.ppf_lookup(s, arrays=False)
    print(f"configs[4]: ms_vote {q.ms_vote:.1f} pairs/s {q.num_scene_pairs / q.ms_vote * 1e3:.3e} votes {q.num_nonunique_votes} K {q.num_top_votes} err {np.linalg.norm(q.pose[:3, 3] - T[:3, 3]):.2f}")
