import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lattice_cryptography_b200 import Engine, make_scheme
secpar = int(sys.argv[1]); n = 1 << int(sys.argv[2])
P = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20), 256: dict(q=39937, l=23, sk_bd=65, ch_wt=50)}[secpar]
eng = Engine(secpar, P['q'], 256, P['l']); eng.use_torch_stream()
sch = make_scheme(sk_bd=P['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=P['ch_wt'])
kc, _ = eng.hash2polyvec('KEY_CH_SEED', ['x'], P['q'] // 2, 256, P['l']); eng.set_key_ch(np.ascontiguousarray(kc[0]))
rng = np.random.default_rng(1)
seeds = (rng.integers(0, 2, (n, secpar), dtype=np.uint8) + 48).astype(np.uint8)
off = (np.arange(n + 1, dtype=np.int64) * secpar)
ds, do = torch.from_numpy(seeds).cuda().view(-1), torch.from_numpy(off).cuda()
for chunk in sys.argv[3:]:
    os.environ['LCB_KEYGEN_CHUNK'] = chunk
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = eng.lm_keygen(sch, (ds, do), want_sk_coef=False, want_vk_coef=False, device=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'secpar {secpar} n {n} chunk {chunk}: {dt*1e3:.1f} ms  {n/dt/1e6:.3f} Mkeys/s')
        del out
