"""Multi-GPU host logic on CPU: world_size-2 gloo.  Each rank votes for its shard of the scene
reference points (the CPU oracle stands in for the CUDA vote stage, which needs a GPU), then the
product's merge (objective_slam_b200.dist.merge_survivors: all_reduce MAX + all_gather) must
reproduce the single-rank survivor list exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from objective_slam_b200.dist import merge_survivors, shard_reference_points
    from oracle import cpu
    g = golden(name)
    df = int(g["ref_df"])
    ns = len(g["scene_pts"])
    mine = set(shard_reference_points(ns, df, rank, world))
    # this rank's accumulator cells: the full histogram restricted to its reference points
    codes, counts = g["hist_codes"], g["hist_counts"]
    sel = np.array([(int(c) >> 32) in mine for c in codes])
    lc = torch.from_numpy(codes[sel].astype(np.int64))
    ln = torch.from_numpy(counts[sel].astype(np.int32))
    lmax = int(ln.max()) if len(ln) else 0
    mc, mn, gmax = merge_survivors(lc, ln, lmax, 0.4)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), codes=mc.numpy().astype(np.uint64), counts=mn.numpy(), gmax=gmax)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny_df1", "tiny_df3"])
def test_two_rank_merge_equals_single_rank(name, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = golden(name)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert int(z["gmax"]) == int(g["vote_counts"][0])
        assert (z["codes"] == g["votes"]).all() and (z["counts"] == g["vote_counts"]).all()


def test_shards_partition_the_reference_points():
    from objective_slam_b200.dist import shard_reference_points
    for n, df, w in [(100, 1, 2), (101, 5, 4), (7, 3, 8), (1, 1, 2), (0, 1, 2)]:
        parts = [shard_reference_points(n, df, r, w) for r in range(w)]
        allrefs = sorted(x for p in parts for x in p)
        assert allrefs == (list(range(0, n, df)) if n > 1 else [])
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
