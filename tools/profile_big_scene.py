"""BASELINE configs[4]-like: dense 1M-point scene (room lattice + one object), 2k-point model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
ns, df = int(sys.argv[1]), int(sys.argv[2])
mp, mn = synth.make_model(2000, seed=0xD209)
sp0, sn0, T = synth.make_scene(mp, mn, 20000, seed=0xD20A)
lp, ln = synth.make_lattice_scene(ns - 20000, pitch=1.0)
sp = np.concatenate([sp0, lp + sp0.min(0)]).astype(np.float32); sn = np.concatenate([sn0, ln]).astype(np.float32)
perm = np.random.default_rng(1).permutation(len(sp)); sp, sn = sp[perm], sn[perm]
d = synth.d_dist_for(mp)
m = ppf.Model(mp, mn, d); s = ppf.Scene(sp, sn, d, df)
for i in range(3):
    r = m.ppf_lookup(s, arrays=False)
    print(f"scene {len(sp)} df {df}: pairs {r.num_scene_pairs:.3e} votes {r.num_nonunique_votes:.3e} ms_vote {r.ms_vote:.2f} pairs/s {r.num_scene_pairs / r.ms_vote * 1e3:.3e} K {r.num_top_votes} err {np.linalg.norm(r.pose[:3,3]-T[:3,3]):.2f}")
