"""Host-visible latency of ONE drop-in keygen (k_sampler_coop + k_matvec), back to back and interleaved with sign/verify
(the pattern bench.py's single_ops uses).  usage: single_keygen_timing.py [reps]"""
import os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_cryptography_b200 import lm_one_time_sigs as lm
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for secpar in (128, 256):
    pp = lm.make_setup_parameters(secpar)
    key = lm.keygen(pp=pp, num_keys_to_gen=1)[0]
    a, b = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); lm.keygen(pp=pp, num_keys_to_gen=1); a.append(time.perf_counter() - t0)
    for _ in range(reps):
        t0 = time.perf_counter(); key = lm.keygen(pp=pp, num_keys_to_gen=1)[0]; b.append(time.perf_counter() - t0)
        sig = lm.sign(pp=pp, otk=key, msg='QRL is awesome!')
        assert lm.verify(pp=pp, otvk=key[2], msg='QRL is awesome!', sig=sig)
    print(f'secpar {secpar}: keygen back to back {1e3 * statistics.median(a):.3f} ms, interleaved with sign + verify {1e3 * statistics.median(b):.3f} ms', flush=True)
