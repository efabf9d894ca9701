"""Where does a ppf_registration call spend its time?  (host wall clock around the C-ABI stages)"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import objective_slam_b200 as ppf
from objective_slam_b200 import synth, _capi as C
SEED = 0xD205 + 2
mp, mn = synth.make_model(10000, seed=SEED); sp, sn, T = synth.make_scene(mp, mn, 50000, seed=SEED + 1); d = synth.d_dist_for(mp)
def t(f):
    torch.cuda.synchronize(); a = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - a) * 1e3
for it in range(3):
    m, tm = t(lambda: ppf.Model(mp, mn, d))
    s, ts = t(lambda: ppf.Scene(sp, sn, d, 8))
    lk, tl = t(lambda: ppf.Lookup())
    _, tv = t(lambda: C.check(C.lib.ppf_model_lookup(m._h, s._h, 8, lk._h)))
    _, td = t(lambda: (m.close(), s.close(), lk.close()))
    print(f"model_create {tm:.1f} ms  scene_create {ts:.1f}  lookup_create {tl:.1f}  lookup {tv:.1f}  destroy {td:.1f}")
    _, tr = t(lambda: ppf.ppf_registration([(sp, sn)], [(mp, mn)], [d], 8))
    print(f"ppf_registration {tr:.1f} ms")
