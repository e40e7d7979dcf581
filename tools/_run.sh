set -x
python -m pytest tests/test_gpu_parity.py -q -x -k "golden or cooperative or keygen_seed or lm_batch" 2>&1 | tail -8
python - <<'PY'
import time, statistics, os, sys
sys.path.insert(0, '.')
from lattice_cryptography_b200 import lm_one_time_sigs as lm
for coop in ('0', '1'):
    os.environ['LCB_SAMPLER_COOP'] = coop
    for secpar in (128, 256):
        pp = lm.make_setup_parameters(secpar)
        lm.keygen(pp=pp, num_keys_to_gen=1)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); lm.keygen(pp=pp, num_keys_to_gen=1); ts.append(time.perf_counter() - t0)
        t8 = []
        for _ in range(5):
            t0 = time.perf_counter(); lm.keygen(pp=pp, num_keys_to_gen=8); t8.append(time.perf_counter() - t0)
        print(f'coop={coop} secpar={secpar}: single keygen {1e3*statistics.median(ts):.2f} ms, 8 keys {1e3*statistics.median(t8):.2f} ms', flush=True)
PY
python tools/sign_timing.py 128 20 2>&1 | tail -1
