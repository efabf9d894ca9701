"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA kernels (oracle/_ref, compiled
from /root/reference for sm_100a) on a B200:  gpurun -- python tests/golden/make_golden.py
The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and the CUDA product
(tests/test_parity_gpu.py::test_golden) to outputs of the reference itself."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from objective_slam_b200 import synth
from oracle import refgpu

CASES = {
    # name: (n_model, n_scene, tau_d, ref_df, seed)
    "tiny_df1": (48, 72, 0.10, 1, 11),
    "tiny_df3": (64, 96, 0.08, 3, 12),
}
OUT = os.environ.get("GOLDEN_OUT", os.path.join(os.path.dirname(os.path.dirname(HERE)), "gpurun_out", "golden"))
os.makedirs(OUT, exist_ok=True)
for name, (nm, ns, tau, df, seed) in CASES.items():
    mp, mn = synth.make_model(nm, seed=seed)
    sp, sn, T = synth.make_scene(mp, mn, ns, seed=seed + 100)
    d = synth.d_dist_for(mp, tau)
    m = refgpu.RefModel(mp, mn, d)
    s = refgpu.RefScene(sp, sn, d, df)
    sppf, skeys = s.features()
    mppf = m.features()
    hk, cnt, first, mapp = m.table()
    hc, hn = m.vote_histogram(s)
    r = m.lookup(s)
    # model keys in pair order, recovered from the table (key of pair map[i] is the key of its bucket)
    mkeys = np.zeros(nm * nm, np.uint32)
    mkeys[mapp.astype(np.int64)] = np.repeat(hk, cnt.astype(np.int64))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), model_pts=mp, model_nrm=mn, scene_pts=sp, scene_nrm=sn, truth=T,
        d_dist=np.float32(d), ref_df=np.int32(df), scene_ppf=sppf, scene_keys=skeys, model_ppf=mppf,
        model_keys=mkeys.reshape(nm, nm), hashkeys=hk, counts=cnt, first=first, map=mapp, hist_codes=hc,
        hist_counts=hn, votes=r["votes"], vote_counts=r["counts"], transformations=r["transformations"],
        weighted=r["weighted"], trans=r["trans"], rots=r["rots"], scores=r["scores"], pose=r["pose"],
        K=np.int64(r["K"]), max_idx=np.int64(r["max_idx"]), num_nonunique_votes=np.int64(r["num_nonunique_votes"]),
        num_unique_votes=np.int64(r["num_unique_votes"]))
    print(name, "K", r["K"], "votes", r["num_nonunique_votes"], "cells", len(hc), "U", len(hk),
          os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")
