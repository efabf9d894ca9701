"""Vote-kernel time of ONE shard of the 8-GPU bench workload on one GPU (reference points rank, rank + 8, ... of the
50k-point scene at ref_point_df = 1): is a slow rank a slow GPU or a slow shard?  usage: shard_tail.py rank [world]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import objective_slam_b200 as ppf
from objective_slam_b200 import synth, _capi as C
rank = int(sys.argv[1]); world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SEED = 0xD205 + 2
mp, mn = synth.make_model(10000, seed=SEED)
sp, sn, T = synth.make_scene(mp, mn, 50000, seed=SEED + 1)
d = synth.d_dist_for(mp)
m = ppf.Model(mp, mn, d); s = ppf.Scene(sp, sn, d, 1); lk = ppf.Lookup()
for i in range(3):
    C.check(C.lib.ppf_lookup_vote(m._h, s._h, 1, rank, world, lk._h))
    st = lk.stats()
    print(f"shard {rank}/{world}: votes {st.num_nonunique_votes} ms_vote {st.ms_vote:.3f} votes/s {st.num_nonunique_votes / st.ms_vote * 1e3:.3e}")
