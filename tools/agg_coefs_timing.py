"""k_agg_coefs time for one shard of a BKLM aggregate: N signatures per aggregate (message of 124 N bytes),
`count` streams starting at `first`, one thread per sponge (LCB_AGG_LANES=1) against two lanes per sponge (=2).
A hashlib subsample checks the coefficients of every run.

  python tools/agg_coefs_timing.py 16 8192 57344      # log2 N, count, first
"""
import hashlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lattice_cryptography_b200 import Engine, make_scheme

log2n, count, first = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
total = 1 << log2n
rng = np.random.default_rng(777)
bits = rng.integers(0, 2, (total, 32), dtype=np.uint8) + ord('0')
agmsg = ('[' + ', '.join(f"(<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * i:010x}>, "
                         f"'{bytes(r).decode()}')" for i, r in enumerate(bits)) + ']').encode()
d_msg = torch.from_numpy(np.frombuffer(agmsg, dtype=np.uint8).copy()).cuda()
perms = count * ((len(agmsg) + 12) // 136 + 1)
variants = [(1, 0, 0), (2, 0, 0), (2, 0, 1)]
for lanes, early, prefetch in variants:
    os.environ['LCB_AGG_LANES'] = str(lanes)
    os.environ['LCB_AGG_PREFETCH'] = str(prefetch)
    eng = Engine(128, 11777, 256, 13)
    eng.use_torch_stream()
    sch = make_scheme()
    for rep in range(2):
        eng.profile(True)
        eng.profile_reset()
        pairs = eng.agg_coefs(sch, d_msg, first, count, device=True)
        torch.cuda.synchronize()
        ms, _ = eng.profile_read('agg_coefs')
        print(f'lanes={lanes} early={early} prefetch={prefetch} N=2^{log2n} count={count} first={first}: {ms:.2f} ms, {perms / ms / 1e6:.3f} Gperm/s', flush=True)
    got = pairs.cpu().numpy()[:, 0, :]
    for i in sorted(set([0, 1, count // 3, count - 2, count - 1] + list(rng.integers(0, count, 6)))):
        dg = hashlib.shake_256(b'AG_SALT' + str(first + i).encode() + agmsg).digest(2)
        assert (int(got[i, 0]), int(got[i, 1])) == (dg[0], 1 if dg[1] & 0x80 else -1), (lanes, i)
    print(f'lanes={lanes} early={early} prefetch={prefetch}: hashlib subsample ok')
    eng.close()
