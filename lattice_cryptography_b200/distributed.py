"""Multi-GPU host logic: one process per GPU (torch.distributed, NCCL over NVLink on B200 boxes,
gloo in CPU tests).  The reference's only parallelism is a multiprocessing.Pool over keys
(lm_one_time_sigs.py:100-123); here

  * keygen / sign / verify / adaptor batches are independent units: every rank takes a contiguous
    range (same split rule as the reference's distribute_tasks, lm_one_time_sigs.py:194-215) and
    there is NO data-path collective;
  * BKLM aggregate / aggregate_verify have one real exchange step: each rank reduces its shard of the
    sorted list to an int32 partial sum (l*d words, or d words for the verification side) and the
    partials are summed with ONE reduce to rank 0, which finishes (mod q, centre / compare).
"""
from typing import Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """(start, count) of rank's contiguous share of n items; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def reduce_partial(partial, dst: int = 0):
    """Sum int32 partial sums over all ranks onto rank `dst` (in place); a no-op without a process
    group.  Accepts a torch tensor (CUDA with nccl, CPU with gloo) or a numpy array (copied)."""
    import torch
    import torch.distributed as dist
    t = partial if isinstance(partial, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(partial))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t


def sharded_aggregate(pp, sig_sorted_local, agmsg, first: int):
    """This rank's part of bklm aggregate: partial sum over its shard (device), reduced to rank 0, which
    returns the aggregate signature int16[l,d] (other ranks return None)."""
    import torch.distributed as dist
    from . import bklm_one_time_agg_sigs as bk
    partial = bk.aggregate_shard(pp, sig_sorted_local, agmsg, first, device=True)
    bk._ctx(pp)[0].synchronize()
    partial = reduce_partial(partial)
    if not dist.is_initialized() or dist.get_rank() == 0:
        return bk.aggregate_finish(pp, partial)
    return None


def sharded_aggregate_verify(pp, vk_ntt_sorted_local, chmsgs_sorted_local, agmsg, first: int, ag_sig, total: int):
    """This rank's part of bklm aggregate_verify; rank 0 returns the verdict, the others None."""
    import torch.distributed as dist
    from . import bklm_one_time_agg_sigs as bk
    partial = bk.aggregate_verify_shard(pp, vk_ntt_sorted_local, chmsgs_sorted_local, agmsg, first, device=True)
    bk._ctx(pp)[0].synchronize()
    partial = reduce_partial(partial)
    if not dist.is_initialized() or dist.get_rank() == 0:
        return bk.aggregate_verify_finish(pp, partial, ag_sig, total)
    return None
