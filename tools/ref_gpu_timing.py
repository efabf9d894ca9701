"""Reference CUDA kernels (oracle/_ref, recompiled for sm_100a) vs libppf_b200 on the same GPU and inputs.
Scene ctor + ppf_lookup with a prebuilt model, CUDA-event timed, median of 5."""
import sys, os, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
from oracle import refgpu

for nm, ns, df in [(1000, 1000, 1), (1500, 3000, 2), (2000, 6000, 5)]:
    mp, mn = synth.make_model(nm, seed=0xD207)
    sp, sn, T = synth.make_scene(mp, mn, ns, seed=0xD208)
    d = synth.d_dist_for(mp)
    rm = refgpu.RefModel(mp, mn, d)
    ref_ms = [rm.time_scene_lookup(sp, sn, df) for _ in range(4)][1:]
    m = ppf.Model(mp, mn, d)
    ours = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s = ppf.Scene(sp, sn, d, df)
        r = m.ppf_lookup(s, arrays=False)
        b.record(); torch.cuda.synchronize()
        ours.append(a.elapsed_time(b))
    ours = ours[1:]
    R = (ns + df - 1) // df
    print(f"model {nm} scene {ns} df {df}: pairs {R * ns} votes {r.num_nonunique_votes} | reference kernels "
          f"{statistics.median(ref_ms):.2f} ms ({R * ns / statistics.median(ref_ms) * 1e3:.3e} pairs/s) | ppf_b200 "
          f"{statistics.median(ours):.3f} ms ({R * ns / statistics.median(ours) * 1e3:.3e} pairs/s) | speed-up "
          f"{statistics.median(ref_ms) / statistics.median(ours):.1f}x")
    del rm
