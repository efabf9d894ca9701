"""First-contact probe: run the reference kernels (oracle/_ref) on the B200 on a small case."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from objective_slam_b200 import synth
from oracle import refgpu

mp, mn = synth.make_model(400)
sp, sn, T = synth.make_scene(mp, mn, 800)
d = synth.d_dist_for(mp)
print("d_dist", d)
t0 = time.time()
m = refgpu.RefModel(mp, mn, d)
hk, cnt, first, mapp = m.table()
print("model table: U=%d N=%d maxbucket=%d  (%.2fs)" % (len(hk), len(mapp), cnt.max(), time.time() - t0))
for df in (1, 5):
    s = refgpu.RefScene(sp, sn, d, df)
    r = m.lookup(s)
    print("df", df, {k: r[k] for k in ("K", "num_nonunique_votes", "num_unique_votes", "max_idx")})
    print(" top counts", r["counts"][:8], "scores", r["scores"][:4])
    print(" pose\n", r["pose"], "\n truth\n", T.astype(np.float32))
    ppf, keys = s.features()
    print(" nan feats", np.isnan(ppf[..., 1:]).sum(), "zero keys", (keys == 0).sum())
