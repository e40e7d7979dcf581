"""Warm k_sign / sampler time of lm_sign for 2^k signatures (CUDA events inside the ABI)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lattice_cryptography_b200 import Engine, make_scheme
secpar = int(sys.argv[1]); n = 1 << int(sys.argv[2])
P = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20), 256: dict(q=39937, l=23, sk_bd=65, ch_wt=50)}[secpar]
eng = Engine(secpar, P['q'], 256, P['l']); eng.use_torch_stream()
sch = make_scheme(sk_bd=P['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=P['ch_wt'])
kc, _ = eng.hash2polyvec('KEY_CH_SEED', ['x'], P['q'] // 2, 256, P['l']); eng.set_key_ch(np.ascontiguousarray(kc[0]))
rng = np.random.default_rng(1)
sk = torch.from_numpy(rng.integers(0, P['q'], (n, 2, P['l'], 256), dtype=np.uint16).view(np.int16)).cuda().view(torch.uint16)
msgs = torch.from_numpy(rng.integers(33, 127, (n, 221), dtype=np.uint8)).cuda().view(-1)
off = torch.arange(n + 1, dtype=torch.int64, device='cuda') * 221
for rep in range(3):
    eng.profile(True); eng.profile_reset()
    sig = eng.lm_sign(sch, sk, (msgs, off), device=True)
    torch.cuda.synchronize()
    print('sign ms', eng.profile_read('sign'), 'sampler ms', eng.profile_read('sampler'))
