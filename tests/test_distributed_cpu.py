"""world_size-2 gloo test of the multi-GPU host logic (lattice_cryptography_b200/distributed.py):
contiguous sharding and the single reduce of int32 partial sums.  The per-shard partial sums are
produced here by a plain numpy statement of BKLM aggregation with monomial coefficients
(sum_i s_i * X^k_i * sig_i, bklm_one_time_agg_sigs.py:96), so the test needs no GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rotate_add(sigs, ks, ss, l, d=256):
    acc = np.zeros((l, d), dtype=np.int64)
    for sig, k, s in zip(sigs, ks, ss):
        rolled = np.roll(sig.astype(np.int64), k, axis=1)
        rolled[:, :k] *= -1
        acc += s * rolled
    return acc


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    from lattice_cryptography_b200.distributed import reduce_partial, shard_range
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    sigs = rng.integers(-945, 946, (n, 13, 256)).astype(np.int16)
    ks, ss = rng.integers(0, 256, n), rng.choice([-1, 1], n)
    start, count = shard_range(n, rank, world)
    part = rotate_add(sigs[start:start + count], ks[start:start + count], ss[start:start + count], 13)
    t = reduce_partial(torch.from_numpy((part % 11777).astype(np.int32)))
    if rank == 0:
        whole = rotate_add(sigs, ks, ss, 13) % 11777
        out.put(bool(np.array_equal(t.numpy().astype(np.int64) % 11777, whole)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from lattice_cryptography_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_reduce_of_partial_sums():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True
