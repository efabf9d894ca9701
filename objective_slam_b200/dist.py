"""Multi-GPU recognition: scene reference points sharded over the ranks of one node.

Every scene reference point owns an independent accumulator (the high 32 bits of a vote code are
s_r, model.h:61-63), so voting needs no data-path collective: rank r votes for every world-th
reference point against a replicated model table.  Only two tiny exchanges couple the ranks, both
required by the reference's semantics (model.cu:160-170):
  1. all_reduce(MAX) of the largest accumulator cell -> the global threshold count > thr * max;
  2. all_gather of each rank's surviving (code, count) list (exact parity needs every survivor,
     not a truncated top-k), after which pose computation runs replicated;
  3. clustering of the merged list is sharded too (every rank scores an interleaved slice of the poses against
     all of them) and the score slices are summed with one all_reduce, after which every rank picks the winner.
One process per GPU.  The product path is inside the library (ppf_model_lookup_sharded, NCCL on the library's own
stream); torch.distributed only ships the NCCL unique id.  `merge_survivors` / `lookup_sharded_torch` state the same
protocol with torch collectives (gloo on CPU for the tests).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _capi as C


def shard_reference_points(n_scene: int, ref_df: int, rank: int, world: int):
    """Scene reference points of one rank: every world-th of the points with s_r % ref_df == 0."""
    refs = range(0, n_scene if n_scene > 1 else 0, ref_df)
    return list(refs)[rank::world]


def merge_survivors(codes: torch.Tensor, counts: torch.Tensor, local_max: int, thr: float, group=None):
    """codes int64 (the 64-bit vote codes), counts int32: this rank's candidates (any superset of its
    survivors).  Returns (codes, counts, global_max) of ALL ranks' survivors, identical on every rank,
    ordered (count desc, code asc) like the reference leaves them (model.cu:148-170)."""
    dev = codes.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    gmax = torch.tensor([int(local_max)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    g = int(gmax.item())
    # float compare exactly as the reference: (float)count > thr * (float)max
    min_votecount = torch.tensor(thr, dtype=torch.float32) * torch.tensor(g, dtype=torch.float32)
    keep = counts.to(torch.float32) > min_votecount.to(dev)
    codes, counts = codes[keep].contiguous(), counts[keep].to(torch.int32).contiguous()
    if world > 1:
        n = torch.tensor([codes.numel()], dtype=torch.int64, device=dev)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        ns = [int(x.item()) for x in ns]
        cap = max(max(ns), 1)
        pc = torch.zeros(cap, dtype=torch.int64, device=dev)
        pn = torch.zeros(cap, dtype=torch.int32, device=dev)
        pc[: codes.numel()] = codes
        pn[: counts.numel()] = counts
        gc = [torch.empty_like(pc) for _ in range(world)]
        gn = [torch.empty_like(pn) for _ in range(world)]
        dist.all_gather(gc, pc, group=group)
        dist.all_gather(gn, pn, group=group)
        codes = torch.cat([c[:k] for c, k in zip(gc, ns)])
        counts = torch.cat([c[:k] for c, k in zip(gn, ns)])
    # (count desc, code asc). Codes are compared as unsigned: flip the sign bit for the signed sort.
    if codes.numel():
        key = codes ^ torch.tensor(-0x8000000000000000, dtype=torch.int64, device=dev)
        order = torch.argsort(key, stable=True)
        codes, counts = codes[order], counts[order]
        order = torch.argsort(counts.to(torch.int64), descending=True, stable=True)
        codes, counts = codes[order], counts[order]
    return codes, counts, g


class Comm:
    """The ranks of a sharded recognition (include/ppf_b200.h, ppf_comm_*): NCCL between processes (one GPU each),
    or host threads of one process on one GPU (`Comm.local`, the single-GPU test vehicle)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_torch(cls, group=None) -> "Comm":
        """One library-owned NCCL communicator over the ranks of a torch.distributed group: rank 0 draws the NCCL
        unique id and torch.distributed (any backend) ships its 128 bytes; the collectives of a lookup then run
        inside the library on its own stream, not through torch."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            C.check(C.lib.ppf_comm_unique_id(buf))
        box = [buf.raw]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        h = ctypes.c_void_p()
        C.check(C.lib.ppf_comm_create_nccl(ctypes.create_string_buffer(box[0], 128), rank, world, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def local(cls, world: int):
        """`world` communicators for `world` host threads of this process (all on the current GPU)."""
        arr = (ctypes.c_void_p * world)()
        C.check(C.lib.ppf_comm_create_local(world, arr))
        return [cls(ctypes.c_void_p(arr[i])) for i in range(world)]

    @property
    def rank(self) -> int:
        return C.lib.ppf_comm_rank(self._h)

    @property
    def size(self) -> int:
        return C.lib.ppf_comm_size(self._h)

    def close(self):
        if getattr(self, "_h", None):
            C.lib.ppf_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def lookup_sharded(model, scene, lookup, comm: Comm, arrays: bool = False):
    """Model::ppf_lookup over the ranks of `comm` (ppf_model_lookup_sharded): the same LookupResult on every rank,
    bit-identical to the single-GPU lookup.  Everything -- voting of the rank's reference points, all_reduce(MAX),
    survivor all_gather, ordering, poses, sharded clustering + all_reduce(SUM), argmax -- runs inside the library."""
    rc = C.check(C.lib.ppf_model_lookup_sharded(model._h, scene._h, scene.ref_point_downsample_factor, comm._h, lookup._h),
                 allow=(C.PPF_ERR_NO_VOTES,))
    return lookup.result(rc, arrays)


def lookup_sharded_torch(model, scene, lookup, rank: int, world: int, group=None, arrays: bool = False):
    """The same exchange written out with torch.distributed collectives over the staged C entry points
    (ppf_lookup_vote / finalize / set_survivors / cluster_shard): the readable statement of the protocol, kept as
    a cross-check of the library path (and usable with any torch backend)."""
    df = scene.ref_point_downsample_factor
    C.check(C.lib.ppf_lookup_vote(model._h, scene._h, df, rank, world, lookup._h))
    lmax = ctypes.c_uint32()
    C.check(C.lib.ppf_lookup_local_max(lookup._h, ctypes.byref(lmax)))
    dev = torch.device("cuda", torch.cuda.current_device())
    gmax = torch.tensor([lmax.value], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    g = int(gmax.item())
    C.check(C.lib.ppf_lookup_finalize(model._h, g, lookup._h))        # local filter with the global max
    if world > 1:
        K = ctypes.c_size_t()
        C.check(C.lib.ppf_lookup_survivors(lookup._h, ctypes.byref(K), None, None))
        codes = torch.empty(max(K.value, 1), dtype=torch.int64, device=dev)
        counts = torch.empty(max(K.value, 1), dtype=torch.int32, device=dev)
        C.check(C.lib.ppf_lookup_copy_survivors(lookup._h, codes.data_ptr(), counts.data_ptr()))
        codes, counts, _ = merge_survivors(codes[: K.value], counts[: K.value], g, model.vote_count_threshold, group)
        torch.cuda.synchronize()
        C.check(C.lib.ppf_lookup_set_survivors(lookup._h, codes.data_ptr(), counts.data_ptr(), codes.numel()))
    C.check(C.lib.ppf_lookup_poses(model._h, scene._h, lookup._h))
    if world > 1 and not model.use_averaged_clusters:
        C.check(C.lib.ppf_lookup_cluster_shard(model._h, lookup._h, rank, world))
        scores = torch.zeros(max(codes.numel(), 1), dtype=torch.float32, device=dev)
        C.check(C.lib.ppf_lookup_copy_scores(lookup._h, scores.data_ptr()))
        dist.all_reduce(scores, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.synchronize()
        C.check(C.lib.ppf_lookup_set_scores(lookup._h, scores.data_ptr()))
        C.check(C.lib.ppf_lookup_cluster_finish(lookup._h))
    else:
        C.check(C.lib.ppf_lookup_cluster(model._h, lookup._h))
    return lookup.result(arrays=arrays)
