#!/usr/bin/env python
"""Search for REACHABLE far-cell FNV collisions: a quantised scene feature whose distance
bin lies beyond every model pair (kd >= K_d) but whose 32-bit key equals a model key.  The reference votes for
such a pair (ppf_vote_count_kernel, kernel.cu:480-501: any scene key equal to a model key is a hit);
tests/test_parity_gpu.py::test_far_cell_collision plants such pairs.  Model keys come from the CPU oracle here
(the test re-derives everything from the GPU table).

    python tools/find_far_collision.py [n_model] [seed]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from objective_slam_b200 import synth        # noqa: E402
from oracle import cpu, farcells             # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 105
    mp, mn = synth.make_model(n, seed=seed)
    ext = float((mp.max(0) - mp.min(0)).max())
    for tau in np.arange(0.0500, 0.0900, 0.0001):
        d = float(np.float32(tau) * np.float32(ext))
        ppf, keys = cpu.scene_features(mp, mn, d, 1)
        K_d = int(np.nanmax(ppf[..., 0]) / d + 0.5) + 1
        for cell, key in farcells.far_collisions(keys[keys != 0], K_d, 4 * K_d, d):
            cnt = int((keys == key).sum())
            ok = farcells.plant_pair(cell, d, (5, 5, 5)) is not None
            print(f"tau={tau:.4f} d={d:.6f} K_d={K_d} far cell={cell} key={key:#x} bucket={cnt} reachable={ok}", flush=True)


if __name__ == "__main__":
    main()
