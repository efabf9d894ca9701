"""Model table build time (ppf_model_create, host clouds in) for the named sizes; configs[2] = 5k points."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
for n in (2000, 5000, 10000):
    mp, mn = synth.make_model(n, seed=0xD205 + 3)
    d = synth.d_dist_for(mp)
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t = time.perf_counter()
        m = ppf.Model(mp, mn, d)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
        U = len(m.table()[0]) if it == 0 else U
        m.close()
    print(f"model {n}: {n * n:.3e} pairs, U {U}, build ms {['%.2f' % x for x in ts]}  -> {n * n / (min(ts) * 1e-3):.3e} pairs/s")
