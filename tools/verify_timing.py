"""Warm lm_verify time for 2^k device-resident triples (challenge sampler + k_verify), verdicts compared with the
construction rule (every 64th triple tampered).
usage: verify_timing.py <secpar> <log2 n> [reps]"""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lattice_cryptography_b200 import Engine, make_scheme
secpar = int(sys.argv[1]); n = 1 << int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
P = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20, vf_bd=945), 256: dict(q=39937, l=23, sk_bd=65, ch_wt=50, vf_bd=3315)}[secpar]
eng = Engine(secpar, P['q'], 256, P['l']); eng.use_torch_stream()
sch = make_scheme(sk_bd=P['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=P['ch_wt'])
kc, _ = eng.hash2polyvec('KEY_CH_SEED', ['x'], P['q'] // 2, 256, P['l']); eng.set_key_ch(np.ascontiguousarray(kc[0]))
rng = np.random.default_rng(1)
seeds = torch.from_numpy(rng.integers(48, 50, (n, secpar), dtype=np.uint8)).cuda()
soff = torch.arange(n + 1, dtype=torch.int64, device='cuda') * secpar
_, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, (seeds.view(-1), soff), want_sk_coef=False, want_vk_coef=False, device=True)
mlen = 93 + secpar
msgs = torch.from_numpy(rng.integers(48, 50, (n, mlen), dtype=np.uint8)).cuda()
off = torch.arange(n + 1, dtype=torch.int64, device='cuda') * mlen
sig = eng.lm_sign(sch, sk_ntt, (msgs.view(-1), off), device=True)
bad = torch.arange(0, n, 64, device='cuda')
sig.view(n, -1)[bad, 7] += 1
expect = torch.ones(n, dtype=torch.uint8, device='cuda'); expect[bad] = 0
out = {}
for rnd in range(2):
    verdict = torch.zeros(n, dtype=torch.uint8, device='cuda')
    eng.lm_verify(sch, vk_ntt, (msgs.view(-1), off), sig, P['vf_bd'], 256, out=verdict)
    torch.cuda.synchronize()
    ok = bool(torch.equal(verdict, expect))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile(True); eng.profile_reset()
    ev0.record()
    for _ in range(reps):
        eng.lm_verify(sch, vk_ntt, (msgs.view(-1), off), sig, P['vf_bd'], 256, out=verdict)
    ev1.record(); torch.cuda.synchronize()
    prof = {k: eng.profile_read(k) for k in ('verify', 'sampler')}
    eng.profile(False)
    print(f'secpar={secpar} n=2^{sys.argv[2]}: {ev0.elapsed_time(ev1) / reps:.3f} ms per step, verdicts as constructed: {ok}, '
          f'{ {k: (round(v[0] / max(v[1], 1), 3), v[1]) for k, v in prof.items()} }', flush=True)
