"""Wall-clock split of one bench step (Scene ctor + sharded lookup) into its C-ABI calls."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import objective_slam_b200 as ppf
from objective_slam_b200 import _capi as C, synth
import bench
mp, mn, sp, sn, d, T = bench.make_workload()
dev = torch.device("cuda", 0)
sp_d, sn_d = torch.from_numpy(sp).to(dev), torch.from_numpy(sn).to(dev)
model = ppf.Model(mp, mn, d); lk = ppf.Lookup(); df = int(os.environ.get("DF", "8"))
def tick(label, fn, acc):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc.setdefault(label, []).append((time.perf_counter() - t) * 1e3); return r
for it in range(int(os.environ.get("ITERS", "6"))):
    acc = {}
    scene = tick("scene", lambda: ppf.Scene(sp_d, sn_d, d, df), acc)
    tick("vote", lambda: C.check(C.lib.ppf_lookup_vote(model._h, scene._h, df, 0, 1, lk._h)), acc)
    lmax = ctypes.c_uint32()
    tick("local_max", lambda: C.check(C.lib.ppf_lookup_local_max(lk._h, ctypes.byref(lmax))), acc)
    tick("finalize", lambda: C.check(C.lib.ppf_lookup_finalize(model._h, lmax.value, lk._h)), acc)
    tick("poses", lambda: C.check(C.lib.ppf_lookup_poses(model._h, scene._h, lk._h)), acc)
    tick("cluster", lambda: C.check(C.lib.ppf_lookup_cluster(model._h, lk._h)), acc)
    res = tick("result", lambda: lk.result(arrays=False), acc)
    tick("close", lambda: scene.close(), acc)
    print(it, {k: round(v[0], 2) for k, v in acc.items()}, "ms_vote", round(res.ms_vote, 1), "K", res.num_top_votes)
