"""Two-rank NCCL test of the sharded BKLM path on real GPUs (skipped on boxes with one GPU; the host
logic of the same path is covered on CPU by tests/test_distributed_cpu.py): every rank keeps its shard
of the sorted list on its own device, partial sums meet in ONE reduce, rank 0 finishes.  The result
must equal the single-GPU aggregate bit for bit and verify."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200 import lattice_algebra as gla
    from lattice_cryptography_b200.distributed import shard_range, sharded_aggregate, sharded_aggregate_verify
    from lattice_cryptography_b200.lattice_algebra import PolynomialVector
    from lattice_cryptography_b200.lm_one_time_sigs import challenge_messages, keygen_batch, sign_batch
    from lattice_cryptography_b200.one_time_keys import SchemeParameters
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    gla.set_default_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    n = 37
    pp = bk.set_aggregation_capacity(bk.make_setup_parameters(128), n)
    lp = pp['scheme_parameters'].lp
    # the same public row on every rank
    kc, _ = gla.engine_for(lp, 128).hash2polyvec('KEY_CH_SEED', ['two-rank test'], lp.modulus // 2, 256, lp.length)
    pp['scheme_parameters'] = SchemeParameters(secpar=128, lp=lp, distribution=pp['scheme_parameters'].distribution,
                                               key_ch=PolynomialVector(lp, _coef=kc[0]))
    seeds = [bin(1000003 * (i + 1))[2:].zfill(128) for i in range(n)]
    ident = [f'<key object at 0x7f{16 * i:08x}>' for i in range(n)]          # already in sorted order
    msgs = [bin(977 * (i + 3))[2:].zfill(32) for i in range(n)]
    agmsg = str(list(zip(ident, msgs)))
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    start, count = shard_range(n, rank, world)
    keys = keygen_batch(pp, seeds[start:start + count], want_coef=False)
    sigs = sign_batch(pp, keys['sk_ntt'], chm[start:start + count])
    ag = sharded_aggregate(pp, sigs, agmsg, start)
    ag_t = torch.from_numpy(ag).cuda() if rank == 0 else torch.empty((lp.length, 256), dtype=torch.int16, device='cuda')
    dist.broadcast(ag_t.view(torch.uint8), src=0)
    ag = ag_t.cpu().numpy()
    ok = sharded_aggregate_verify(pp, keys['vk_ntt'], chm[start:start + count], agmsg, start, ag, n)
    bad = ag.copy()
    bad[3, 7] += 1
    ok_bad = sharded_aggregate_verify(pp, keys['vk_ntt'], chm[start:start + count], agmsg, start, bad, n)
    if rank == 0:
        # single-GPU reference of the same aggregate
        all_keys = keygen_batch(pp, seeds, want_coef=False)
        all_sigs = sign_batch(pp, all_keys['sk_ntt'], chm)
        whole = bk.aggregate_finish(pp, bk.aggregate_shard(pp, all_sigs, agmsg, 0))
        out.put((bool(np.array_equal(whole, ag)), ok, ok_bad))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_bklm():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    same, ok, ok_bad = out.get(timeout=10)
    assert same and ok is True and ok_bad is False
