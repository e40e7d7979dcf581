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


# ------------------------------------------------------------------------------------------------
# The other halves of the metric: keygen / sign phases in the shape of
# benchmarks/benchmark_lm_one_time_sigs.py:35-135, BKLM aggregate / aggregate-verify at small N with an
# a*N^2 + b*N extrapolation to the benchmark size, and the adaptor flow of tests/test_adaptor_sigs.py:196-217.
def _pool_map(fn, jobs):
    if len(jobs) == 1:
        return [fn(jobs[0])]
    with mp.get_context('fork').Pool(len(jobs)) as pool:
        return pool.map(fn, jobs)


def _split(sample: int, nproc: int):
    nproc = max(1, min(nproc or os.cpu_count() or 1, sample))
    bounds = [(sample * i) // nproc for i in range(nproc + 1)]
    return [(a, b - a) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]


def _lm_phase_worker(args):
    secpar, first, count = args
    import schemes
    pp = schemes.make_lm_parameters(secpar, schemes.key_ch_from_seed(secpar, KEY_CH_SEED))
    t = {}
    t0 = time.perf_counter()
    keys = [schemes.lm_keygen_one(pp, bench_seed(secpar, i)) for i in range(first, first + count)]
    t['keygen'] = time.perf_counter() - t0
    msgs = [bench_chmsg(secpar, i) for i in range(first, first + count)]
    t0 = time.perf_counter()
    sigs = [schemes.lm_sign(pp, k[0], k[1], m) for k, m in zip(keys, msgs)]
    t['sign'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ok = [schemes.lm_verify(pp, k[2], k[3], m, s) for k, m, s in zip(keys, msgs, sigs)]
    t['verify'] = time.perf_counter() - t0
    return t, all(ok)


def lm_phase_throughput(secpar: int, sample: int, nproc: int = None) -> dict:
    """Whole-host keygen / sign / verify rates of the CPU port on `sample` seeds (each phase timed per worker,
    rate = sample / slowest worker)."""
    jobs = [(secpar, a, c) for a, c in _split(sample, nproc)]
    res = _pool_map(_lm_phase_worker, jobs)
    if not all(r[1] for r in res):
        raise RuntimeError('CPU port rejected an honest signature')
    out = {'cores': len(jobs), 'sample': sample, 'secpar': secpar}
    for ph in ('keygen', 'sign', 'verify'):
        worst = max(r[0][ph] for r in res)
        out[f'{ph}_per_s'] = sample / worst
        out[f'{ph}_ms_per_op_per_core'] = 1e3 * sum(r[0][ph] for r in res) / sample
    return out


def _bklm_material(secpar: int, n: int):
    """n keys / 32-bit messages / signatures and the aggregation message, deterministic."""
    import schemes
    pp = schemes.make_bklm_parameters(secpar, schemes.key_ch_from_seed(secpar, KEY_CH_SEED), n)
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * i:010x}>' for i in range(n)]
    msgs = [bin((0xD1B54A32D192ED03 * (i + 7)) % (1 << 32))[2:].zfill(32) for i in range(n)]
    agmsg = '[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']'
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    return pp, agmsg, chm


def _bklm_make_worker(args):
    secpar, n, first, count = args
    import schemes
    pp, agmsg, chm = _bklm_material(secpar, n)
    out = []
    for i in range(first, first + count):
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, bench_seed(secpar, i))
        out.append((schemes.dense_of_poly(vkl), schemes.dense_of_poly(vkr),
                    schemes.dense_of_vec(schemes.lm_sign(pp, skl, skr, chm[i]))))
    return out


def _bklm_agg_worker(args):
    """One worker's share of sum_i sig_i ** ag_i (bklm_one_time_agg_sigs.py:96)."""
    secpar, n, first, count, sigs = args
    import schemes
    pp, agmsg, _ = _bklm_material(secpar, n)
    lp = pp['lp']
    svecs = [schemes.vec_from_dense(lp, s) for s in sigs]
    t0 = time.perf_counter()
    coefs = [schemes.sample_poly(secpar, lp, pp['ag_salt'] + str(i), agmsg, pp['ag_bd'], pp['ag_wt'])
             for i in range(first, first + count)]
    part = sum([s ** a for s, a in zip(svecs, coefs)])
    return time.perf_counter() - t0, schemes.dense_of_vec(part)


def _bklm_aggv_worker(args):
    """One worker's share of sum_i (vk_left_i * c_i + vk_right_i) * ag_i (bklm_one_time_agg_sigs.py:107-115)."""
    secpar, n, first, count, vks = args
    import schemes
    pp, agmsg, chm = _bklm_material(secpar, n)
    lp = pp['lp']
    vk = [(schemes.poly_from_dense(lp, a), schemes.poly_from_dense(lp, b)) for a, b in vks]
    t0 = time.perf_counter()
    challs = [schemes.challenge(pp, chm[i]) for i in range(first, first + count)]
    coefs = [schemes.sample_poly(secpar, lp, pp['ag_salt'] + str(i), agmsg, pp['ag_bd'], pp['ag_wt'])
             for i in range(first, first + count)]
    part = sum([(v[0] * c + v[1]) * a for a, c, v in zip(coefs, challs, vk)])
    return time.perf_counter() - t0, schemes.dense_of_poly(part)


def bklm_times(secpar: int, n: int, nproc: int = None) -> dict:
    """Wall time of ONE aggregate and ONE aggregate-verify of n signatures with the sum split over the host cores
    (the reference is single-threaded; this is the best the host can do with its algorithm), checked: the
    aggregate must verify."""
    import schemes
    jobs = _split(n, nproc)
    made = [x for r in _pool_map(_bklm_make_worker, [(secpar, n, a, c) for a, c in jobs]) for x in r]
    pp, agmsg, chm = _bklm_material(secpar, n)
    lp = pp['lp']
    res = _pool_map(_bklm_agg_worker, [(secpar, n, a, c, [m[2] for m in made[a:a + c]]) for a, c in jobs])
    t0 = time.perf_counter()
    ag_sig = sum([schemes.vec_from_dense(lp, r[1]) for r in res])
    t_agg = max(r[0] for r in res) + (time.perf_counter() - t0)
    resv = _pool_map(_bklm_aggv_worker, [(secpar, n, a, c, [(m[0], m[1]) for m in made[a:a + c]]) for a, c in jobs])
    t0 = time.perf_counter()
    cnw = ag_sig.get_coef_rep()
    nn, w = max(i[1] for i in cnw), max(i[2] for i in cnw)
    ok = 1 <= nn <= pp['avf_bd'] and 1 <= w <= pp['avf_wt'] and 1 <= n <= pp['ag_cap']
    total = sum([schemes.poly_from_dense(lp, r[1]) for r in resv])
    ok = ok and (pp['key_ch'] * ag_sig == total)
    t_aggv = max(r[0] for r in resv) + (time.perf_counter() - t0)
    if not ok:
        raise RuntimeError(f'CPU port: aggregate of {n} signatures did not verify')
    return {'n': n, 'aggregate_s': t_agg, 'aggregate_verify_s': t_aggv, 'cores': len(jobs)}


def _hash_rate_worker(args):
    nbytes, reps = args
    import hashlib
    msg = bytes(nbytes)
    t0 = time.perf_counter()
    for i in range(reps):
        hashlib.shake_256(b'AG_SALT' + str(i).encode() + msg).digest(2)
    return nbytes * reps / (time.perf_counter() - t0)


def bklm_extrapolation(secpar: int, ns=(2, 16, 64, 256), target: int = 1 << 16, nproc: int = None) -> dict:
    """t(N) = a*N^2 + b*N per aggregate.  b comes from a least-squares fit through the
    measured points (b*N is the per-signature algebra and challenge); the quadratic term is the N hashes of a
    124*N-byte message (bklm_one_time_agg_sigs.py:60-81), invisible at the N a bounded run can afford, so `a` is
    taken from hashlib's SHAKE256 rate measured on every core AT the target message size:
    a = 124 / (bytes per second per core * cores)."""
    cores = max(1, nproc or os.cpu_count() or 1)
    pts = [bklm_times(secpar, n, cores) for n in ns]
    rate = min(_pool_map(_hash_rate_worker, [(124 * target, 2)] * cores))            # B/s per core, all cores busy
    a = 124.0 / (rate * cores)
    out = {'points': pts, 'cores': cores, 'shake256_bytes_per_s_per_core': rate, 'a_s_per_sig2': a, 'target_n': target}
    for key in ('aggregate', 'aggregate_verify'):
        # least squares for b with a fixed: minimise sum (t - a n^2 - b n)^2
        num = sum((p[f'{key}_s'] - a * p['n'] ** 2) * p['n'] for p in pts)
        b = num / sum(p['n'] ** 2 for p in pts)
        t = a * target ** 2 + b * target
        out[key] = {'b_s_per_sig': b, 'extrapolated_s': t, 'sigs_per_s': target / t,
                    'largest_n_run': pts[-1]['n'], 'largest_n_sigs_per_s': pts[-1]['n'] / pts[-1][f'{key}_s']}
    return out


def _adaptor_worker(args):
    secpar, first, count = args
    import schemes
    pp = schemes.make_adaptor_parameters(secpar, schemes.key_ch_from_seed(secpar, KEY_CH_SEED))
    t = {k: 0.0 for k in ('witgen', 'presign', 'preverify', 'adapt', 'verify', 'extract', 'witness_verify')}
    ok = True
    for i in range(first, first + count):
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, bench_seed(secpar, i))
        m = '<st>, ' + bench_chmsg(secpar, i)
        t0 = time.perf_counter(); wit, st = schemes.witgen_one(pp, bench_seed(secpar, i + (1 << 20))); t['witgen'] += time.perf_counter() - t0
        t0 = time.perf_counter(); presig = schemes.lm_sign(pp, skl, skr, m); t['presign'] += time.perf_counter() - t0
        t0 = time.perf_counter(); ok &= schemes.lm_verify(pp, vkl, vkr, m, presig, None, 'pvf_bd', 'pvf_wt'); t['preverify'] += time.perf_counter() - t0
        t0 = time.perf_counter(); sig = schemes.adapt(presig, wit); t['adapt'] += time.perf_counter() - t0
        t0 = time.perf_counter(); ok &= schemes.lm_verify(pp, vkl, vkr, m, sig, st); t['verify'] += time.perf_counter() - t0
        t0 = time.perf_counter(); ext = schemes.extract(presig, sig); t['extract'] += time.perf_counter() - t0
        t0 = time.perf_counter(); ok &= schemes.witness_verify(pp, ext, st); t['witness_verify'] += time.perf_counter() - t0
    return t, ok


def adaptor_throughput(secpar: int, instances: int = 64, nproc: int = None) -> dict:
    jobs = [(secpar, a, c) for a, c in _split(instances, nproc)]
    res = _pool_map(_adaptor_worker, jobs)
    if not all(r[1] for r in res):
        raise RuntimeError('CPU port: adaptor flow failed')
    return {'cores': len(jobs), 'instances': instances,
            'ops_per_s': {k: instances / max(r[0][k] for r in res) for k in res[0][0]}}
