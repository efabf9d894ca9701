"""Far-cell FNV collisions -- TEST INFRASTRUCTURE ONLY (numpy restatement of ppf_hash_kernel, kernel.cu:460-477,
over the quantised-feature lattice).

The reference matches a scene pair to a model bucket by equality of the 32-bit FNV key alone
(ppf_vote_count_kernel, kernel.cu:480-501).  A scene pair whose distance bin lies beyond every model pair
(kd >= K_d) therefore still votes when its key collides with a model key.  These helpers enumerate such cells and
construct an oriented point pair that lands in a given cell, for tests/test_parity_gpu.py::test_far_cell_collision
and tools/find_far_collision.py.
"""
import numpy as np

D_ANGLE = np.float32(np.float32(2.0) * np.float32(3.14159274101257324) / np.float32(30.0))   # kernel.h:16


def fnv_cells(kd, k1, k2, k3, d_dist):
    """FNV-1a (signed bytes, kernel.cu:23-30) of the float4 (RN(kd d), RN(k1 D), RN(k2 D), RN(k3 D)); k = 16 means NaN
    (canonical 0x7FFFFFFF, what acosf / fmodf leave)."""
    def bits(k, step):
        return (k.astype(np.float32) * np.float32(step)).astype(np.float32).view(np.uint32)
    words = [bits(kd, d_dist)]
    for k in (k1, k2, k3):
        w = bits(np.minimum(k, 15), D_ANGLE)
        words.append(np.where(k >= 16, np.uint32(0x7FFFFFFF), w))
    h = np.full(kd.shape, 2166136261, np.uint32)
    for w in words:
        for b in range(4):
            byte = ((w >> np.uint32(8 * b)) & np.uint32(0xFF)).astype(np.uint8).view(np.int8).astype(np.int32).astype(np.uint32)
            h = h ^ byte
            h = (h.astype(np.uint64) * np.uint64(16777619) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return h


def far_collisions(model_keys, K_d, K_far, d_dist):
    """Cells (kd, k1, k2, k3) with K_d <= kd < K_far whose key is one of `model_keys` (non-zero): list of
    ((kd, k1, k2, k3), key)."""
    uk = np.unique(np.asarray(model_keys, np.uint32))
    uk = uk[uk != 0]
    out = []
    for kd0 in range(K_d, K_far, 64):
        kd, k1, k2, k3 = np.meshgrid(np.arange(kd0, min(kd0 + 64, K_far)), np.arange(17), np.arange(17), np.arange(17),
                                     indexing="ij")
        kd, k1, k2, k3 = [x.ravel() for x in (kd, k1, k2, k3)]
        h = fnv_cells(kd, k1, k2, k3, d_dist)
        for i in np.nonzero(np.isin(h, uk))[0]:
            out.append(((int(kd[i]), int(k1[i]), int(k2[i]), int(k3[i])), int(h[i])))
    return out


def angles_for_cell(k1, k2, k3):
    """(theta1, theta2, theta3) inside the three angle bins that a pair of oriented points can realise
    (|t1 - t2| <= t3 <= min(t1 + t2, 2 pi - t1 - t2)), as far from the bin edges as possible; None if there is
    none.  k = 16 (NaN angle) is only handled for the zero-normal patterns (16, x, 16) and (x, 16, 16)."""
    D = float(D_ANGLE)
    best, best_m = None, 0.0
    g = 13
    rng = [np.linspace(k * D, min((k + 1) * D, np.pi), g + 2)[1:-1] for k in (min(k1, 15), min(k2, 15), min(k3, 15))]
    for a in rng[0]:
        for b in rng[1]:
            for c in rng[2]:
                lo, hi = abs(a - b), min(a + b, 2 * np.pi - a - b)
                m = min(c - lo, hi - c,
                        a - k1 * D, min((k1 + 1) * D, np.pi) - a, b - k2 * D, min((k2 + 1) * D, np.pi) - b,
                        c - k3 * D, min((k3 + 1) * D, np.pi) - c)
                if m > best_m:
                    best, best_m = (a, b, c), m
    return best if best_m > 2e-3 else None


def plant_pair(cell, d_dist, origin):
    """Two oriented points (p1, n1, p2, n2) whose pair feature (p1 -> p2) falls in `cell`; None when the cell is
    not reachable.  d = p2 - p1 runs along +x."""
    kd, k1, k2, k3 = cell
    L = (kd + 0.5) * d_dist
    p1 = np.asarray(origin, np.float64)
    p2 = p1 + np.array([L, 0.0, 0.0])
    if k1 == 16 and k3 == 16 and k2 < 16:        # n1 = 0: angle(n1, d) and angle(n1, n2) are 0/0 = NaN
        t2 = (k2 + 0.5) * float(D_ANGLE)
        return p1, np.zeros(3), p2, np.array([np.cos(t2), np.sin(t2), 0.0])
    if k2 == 16 and k3 == 16 and k1 < 16:        # n2 = 0
        t1 = (k1 + 0.5) * float(D_ANGLE)
        return p1, np.array([np.cos(t1), np.sin(t1), 0.0]), p2, np.zeros(3)
    if max(k1, k2, k3) >= 16:
        return None
    ang = angles_for_cell(k1, k2, k3)
    if ang is None:
        return None
    t1, t2, t3 = ang
    n1 = np.array([np.cos(t1), np.sin(t1), 0.0])
    cphi = (np.cos(t3) - np.cos(t1) * np.cos(t2)) / (np.sin(t1) * np.sin(t2))
    cphi = float(np.clip(cphi, -1.0, 1.0))
    sphi = np.sqrt(max(0.0, 1.0 - cphi * cphi))
    n2 = np.array([np.cos(t2), np.sin(t2) * cphi, np.sin(t2) * sphi])
    return p1, n1, p2, n2
