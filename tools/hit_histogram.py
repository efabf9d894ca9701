"""How the votes of the bench workload distribute over the multiplicity h of a (reference point, bucket) hit set.

CPU-only statistics (numpy, approximate float64 quantiser: a pair on a bin edge may land in the neighbour cell,
irrelevant for a histogram).  For a sample of scene reference points: h(cell) = number of scene points whose pair with
the reference point falls in the cell, L(cell) = number of model pairs in the cell; votes(cell) = h * L.  Prints the
share of votes cast by hit sets with h >= 1, 2, 4, ... and the number of such hit sets per reference point: the
decision data for a count-based (fewer than one atomic per vote) formulation.
usage: hit_histogram.py [n_model n_scene n_refs]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("synth", os.path.join(os.path.dirname(__file__), "..", "objective_slam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); sys.modules["synth"] = synth; spec.loader.exec_module(synth)

nm = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
nrefs = int(sys.argv[3]) if len(sys.argv) > 3 else 48
SEED = 0xD205 + 2
mp, mn = synth.make_model(nm, seed=SEED)
sp, sn, T = synth.make_scene(mp, mn, ns, seed=SEED + 1)
d_dist = synth.d_dist_for(mp, 0.05)
DA = 2 * np.pi / 30


def cells(p1, n1, P2, N2):
    d = P2 - p1
    dn = np.linalg.norm(d, axis=1)
    with np.errstate(all="ignore"):
        a1 = np.arccos(np.clip((d @ n1) / (dn * np.linalg.norm(n1)), -1, 1))
        a2 = np.arccos(np.clip(np.einsum("ij,ij->i", d, N2) / (dn * np.linalg.norm(N2, axis=1)), -1, 1))
        a3 = np.arccos(np.clip((N2 @ n1) / (np.linalg.norm(N2, axis=1) * np.linalg.norm(n1)), -1, 1))
    kd = np.floor(dn / d_dist).astype(np.int64)
    k = [np.nan_to_num(np.floor(a / DA)).astype(np.int64) for a in (a1, a2, a3)]
    return ((kd * 17 + k[0]) * 17 + k[1]) * 17 + k[2], dn > 0


mp64, mn64, sp64, sn64 = (x.astype(np.float64) for x in (mp, mn, sp, sn))
KD = int(np.ceil(np.linalg.norm(mp64.max(0) - mp64.min(0)) / d_dist)) + 2
L = np.zeros(KD * 17 ** 3, np.int64)
for r in range(nm):
    c, ok = cells(mp64[r], mn64[r], mp64, mn64)
    ok &= np.arange(nm) != r
    np.add.at(L, c[ok], 1)
print(f"model {nm}: {int((L > 0).sum())} occupied cells, max L {int(L.max())}, pairs {int(L.sum())}")

edges = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 1 << 30]
votes_by = np.zeros(len(edges) - 1); sets_by = np.zeros(len(edges) - 1); entries_by = np.zeros(len(edges) - 1)
rng = np.random.default_rng(5)
refs = rng.choice(np.arange(0, ns, 8), nrefs, replace=False)
per_ref = []
for r in refs:
    c, ok = cells(sp64[r], sn64[r], sp64, sn64)
    ok &= (np.arange(ns) != r) & (c < len(L))
    cu, h = np.unique(c[ok], return_counts=True)
    Lc = L[cu]
    hit = Lc > 0
    cu, h, Lc = cu[hit], h[hit], Lc[hit]
    v = h * Lc
    for i in range(len(edges) - 1):
        m = (h >= edges[i]) & (h < edges[i + 1])
        votes_by[i] += v[m].sum(); sets_by[i] += m.sum(); entries_by[i] += Lc[m].sum()
    per_ref.append((int(v.sum()), int((h >= 64).sum()), int(v[h >= 64].sum()), int(h.sum())))
tot = votes_by.sum()
print(f"{nrefs} reference points: {tot / nrefs:.3e} votes per reference point, hits per ref {np.mean([p[3] for p in per_ref]):.0f}")
print(" h in [a,b)      share of votes   cum share(h>=a)   hit sets/ref   entries/ref (L summed)")
cum = 1.0
for i in range(len(edges) - 1):
    print(f" [{edges[i]:5d},{edges[i+1] if edges[i+1] < 1 << 29 else 'inf':>5})   {votes_by[i] / tot:12.4f}   {cum:12.4f}   {sets_by[i] / nrefs:12.1f}   {entries_by[i] / nrefs:14.3e}")
    cum -= votes_by[i] / tot
pr = np.array(per_ref)
print("per reference point: votes min/median/max %.2e %.2e %.2e; heavy (h>=64) sets per ref min/median/max %d %d %d" % (
    pr[:, 0].min(), np.median(pr[:, 0]), pr[:, 0].max(), pr[:, 1].min(), np.median(pr[:, 1]), pr[:, 1].max()))
