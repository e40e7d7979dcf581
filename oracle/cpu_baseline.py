"""CPU ORACLE (test infrastructure): times the reference's CPU algorithm for the hot path on
the host cores.  Used ONLY by bench.py's `cpu_baseline` leg and `--impl reference` arm.

What runs: oracle/schemes.py (the restated scheme layer) on oracle/lattice_algebra (the restated
pure-Python L1: 2d-point cyclic transform on Python integers, i.e. the reference's cost
structure).  The reference's own modules cannot travel to the GPU box (/root/reference is absent
there), so kind = "port".  As in benchmarks/benchmark_lm_one_time_sigs.py:105-135, keys,
messages and signatures are produced BEFORE the clock starts (the reference's verify receives
Polynomial objects that are already in NTT form); only verify() calls are timed.  Workers are
separate processes (the reference is GIL-bound pure Python), one per host core, each timing its
own share; whole-host throughput = sample / slowest worker.
"""
import multiprocessing as mp
import os
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

KEY_CH_SEED = 'lcb200 bench key_ch v1'


def bench_seed(secpar: int, i: int) -> str:
    return bin((0x9E3779B97F4A7C15 * (i + 1)) % (1 << secpar))[2:].zfill(secpar)


def bench_chmsg(secpar: int, i: int) -> str:
    """'<...OneTimeVerificationKey object at 0x7f..........>, ' + secpar-bit message: the shape of
    str(otvk)+', '+msg in lm_one_time_sigs.py:148 with benchmark_lm_one_time_sigs.py:75 messages."""
    msg = bin((0xD1B54A32D192ED03 * (i + 7)) % (1 << secpar))[2:].zfill(secpar)
    return f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * i:010x}>, {msg}'


def _lm_verify_worker(args):
    secpar, first, count = args
    import schemes
    pp = schemes.make_lm_parameters(secpar, schemes.key_ch_from_seed(secpar, KEY_CH_SEED))
    items = []
    for i in range(first, first + count):
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, bench_seed(secpar, i))
        m = bench_chmsg(secpar, i)
        items.append((vkl, vkr, m, schemes.lm_sign(pp, skl, skr, m)))
    t0 = time.perf_counter()
    verdicts = [schemes.lm_verify(pp, vkl, vkr, m, sig) for vkl, vkr, m, sig in items]
    return time.perf_counter() - t0, verdicts


def lm_verify_throughput(secpar: int, sample: int, nproc: int = None) -> dict:
    """Whole-host LM-OTS verifies/s of the CPU port on `sample` honest triples."""
    nproc = max(1, min(nproc or os.cpu_count() or 1, sample))
    bounds = [(sample * i) // nproc for i in range(nproc + 1)]
    jobs = [(secpar, a, b - a) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    if len(jobs) == 1:
        results = [_lm_verify_worker(jobs[0])]
    else:
        with mp.get_context('fork').Pool(len(jobs)) as pool:
            results = pool.map(_lm_verify_worker, jobs)
    elapsed = max(r[0] for r in results)
    verdicts = [v for r in results for v in r[1]]
    if not all(verdicts):
        raise RuntimeError('CPU port rejected an honest signature')
    return {'value': sample / elapsed, 'cores': len(jobs), 'elapsed_s': elapsed, 'sample': sample}
