#!/usr/bin/env python
"""bench.py — headline benchmark of the lcb200 engine.

Metric (BASELINE.json): LM-OTS verifies/sec.  Workload (BASELINE.json configs[1]): batched LM-OTS
verify of 2^20 independent (vk, msg, sig) triples at secpar = 128 per B200.  A "step" is one pass
of the verify path (challenge sampler kernel + verify kernel) over the whole batch.  With N > 1 GPUs
every rank verifies its own 2^20 triples (independent units, no data-path collective): weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # engine arm
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU algorithm (oracle port)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHIPPED = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20, vf_bd=945, vf_wt=256),
           256: dict(q=39937, l=23, sk_bd=65, ch_wt=50, vf_bd=3315, vf_wt=256)}
D = 256
KEY_CH_SEED = 'lcb200 bench key_ch v1'
METRIC = 'LM-OTS verifies/sec'
UNIT = 'verifies/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--secpar', type=int, default=128, choices=[128, 256])
    ap.add_argument('--log2n', type=int, default=20, help='log2 of triples per GPU')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--cpu-per-core', type=int, default=32, help='CPU baseline: verifies per host core')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the keygen+sign / BKLM / adaptor measurements')
    ap.add_argument('--bklm-log2n', type=int, default=16, help='log2 of signatures per aggregate (whole job)')
    return ap.parse_args()


def workload_name(a):
    return f'lm_ots_verify_batch secpar={a.secpar} n=2^{a.log2n} triples per GPU'


# ---------------------------------------------------------------------------------------------------
def cpu_leg(secpar, per_core):
    """The oracle port on every host core, bounded sample (see oracle/cpu_baseline.py)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpu_baseline
    cores = os.cpu_count() or 1
    sample = per_core * cores
    r = cpu_baseline.lm_verify_throughput(secpar, sample, cores)
    return {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
            'sample': f'{sample} honest LM-OTS triples at secpar {secpar} ({per_core} per core), verify() only timed, '
                      f'pure-Python port of lattice_algebra (oracle/), slowest worker {r["elapsed_s"]:.2f} s'}


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals, t0 = [], time.perf_counter()
    leg = None
    for step in range(a.warmup + a.steps):
        leg = cpu_leg(a.secpar, max(1, a.cpu_per_core // 2))
        if step >= a.warmup:
            vals.append(leg['value'])
    v = sum(vals) / len(vals)
    leg['value'] = v
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': 1e3 * (time.perf_counter() - t0) / (a.warmup + a.steps),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'python-int', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'l2': 'n/a (CPU)'},
            'cpu_baseline': leg,
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is
    allocated: the end-to-end legs stream 6-8 GB per step from pinned memory, and first-touch placement on the
    GPU's own NUMA node keeps that traffic off the inter-socket link.  Best effort; returns what it did."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(torch.cuda.get_device_properties(local).uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler(object):
    """SM clock / power / throttle reasons sampled every 5 ms by NVML in a thread DURING the timed region."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, device):
        import threading
        self.rows, self.stop_flag, self.h = [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:        # by UUID, so that CUDA_VISIBLE_DEVICES remapping cannot point NVML at another GPU
                import torch
                self.h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(torch.cuda.get_device_properties(device).uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0, reasons))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.h is None:
            return out
        self.stop_flag = True
        self.t.join(timeout=2)
        if not self.rows:
            return out
        sm = sorted(r[0] for r in self.rows)
        out['sm_mhz'] = float(sm[len(sm) // 2])
        out['sm_min_mhz'] = float(sm[0])
        out['sm_max_mhz'] = float(self.max_sm)
        out['power_w_max'] = max(r[1] for r in self.rows)
        seen = 0
        for r in self.rows:
            seen |= r[2]
        out['reasons'] = [name for bit, name in self.REASONS.items() if seen & bit]
        out['samples'] = len(self.rows)
        return out


def build_inputs(a, rank, torch, np):
    """Synthetic seeds / messages of the benchmark's shape, built as byte matrices (no Python loops)."""
    n = 1 << a.log2n
    rng = np.random.default_rng(20260101 + rank)
    seeds = (rng.integers(0, 2, (n, a.secpar), dtype=np.uint8) + ord('0')).astype(np.uint8)
    head = np.frombuffer(b'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f', dtype=np.uint8)
    idx = (np.arange(n, dtype=np.int64) + rank * n) * 16
    hexd = np.frombuffer(b'0123456789abcdef', dtype=np.uint8)
    addr = np.stack([hexd[(idx >> (4 * (9 - k))) & 15] for k in range(10)], axis=1)
    tail = np.frombuffer(b'>, ', dtype=np.uint8)
    bits = (rng.integers(0, 2, (n, a.secpar), dtype=np.uint8) + ord('0')).astype(np.uint8)
    ch = np.concatenate([np.broadcast_to(head, (n, head.size)), addr, np.broadcast_to(tail, (n, 3)), bits], axis=1)
    ch = np.ascontiguousarray(ch)
    seed_off = np.arange(n + 1, dtype=np.int64) * a.secpar
    ch_off = np.arange(n + 1, dtype=np.int64) * ch.shape[1]
    return seeds.reshape(-1), seed_off, ch, ch_off



# ---------------------------------------------------------------------------------------------------
# Secondary measurements (BASELINE.json configs[2..4] at sizes that keep the default run short).  Timed
# with CUDA events on the engine's stream, max over ranks; none of them is the headline metric.
def _timed(torch, dist, world, dev, fn, reps=3):
    """Median device time of `reps` calls after one warm-up call (max over ranks).  The median keeps a one-off
    allocator hiccup (torch or the stream-ordered pool growing) out of the secondary figures."""
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = torch.tensor([times[len(times) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _engine(secpar, local, np):
    from lattice_cryptography_b200 import Engine, make_scheme
    p = SHIPPED[secpar]
    eng = Engine(secpar, p['q'], D, p['l'], device=local)
    eng.use_torch_stream()
    key_ch, _ = eng.hash2polyvec('KEY_CH_SEED', [KEY_CH_SEED], p['q'] // 2, D, p['l'])
    eng.set_key_ch(np.ascontiguousarray(key_ch[0]))
    sch = make_scheme(sk_bd=p['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=p['ch_wt'], wit_bd=1, wit_wt=20)
    return eng, sch, p


def secondary_keygen_sign(a, rank, local, world, torch, np, dist):
    """configs[2]: keygen + sign from synthetic seeds at secpar 256; 2^17 seeds per GPU (the config's 2^20
    seeds over 8 GPUs)."""
    class A: secpar, log2n = 256, 17
    eng, sch, p = _engine(256, local, np)
    dev = f'cuda:{local}'
    n = 1 << A.log2n
    seeds, seed_off, ch, ch_off = build_inputs(A, rank, torch, np)
    d_seeds = (torch.from_numpy(seeds).to(dev), torch.from_numpy(seed_off).to(dev))
    d_ch = (torch.from_numpy(ch).to(dev).view(-1), torch.from_numpy(ch_off).to(dev))
    out = {}

    def run():
        _, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, d_seeds, want_sk_coef=False, want_vk_coef=False, device=True)
        out['sig'] = eng.lm_sign(sch, sk_ntt, d_ch, device=True)
        out['vk'] = vk_ntt
    ms = _timed(torch, dist, world, dev, run, reps=3)
    ok = bool(eng.lm_verify(sch, out['vk'], d_ch, out['sig'], p['vf_bd'], p['vf_wt'], device=True).all().item())
    eng.close()
    perms = 2 * 2853 + 26
    return {'keys_per_s': world * n / (ms * 1e-3), 'ms': ms, 'n_per_gpu': n, 'all_verify': ok,
            'keccak_gperm_s_per_gpu': n * perms / (ms * 1e-3) / 1e9, 'keccak_roofline_gperm_s': 4.28}


def secondary_verify_secpar256(a, rank, local, world, torch, np, dist):
    """The headline path on the other shipped parameter set: 2^18 (vk, msg, sig) triples per GPU at secpar 256
    (q = 39937, l = 23, ch_wt = 50; 13.2 KB and 26 permutations per verify)."""
    class A: secpar, log2n = 256, 18
    eng, sch, p = _engine(256, local, np)
    dev = f'cuda:{local}'
    n = 1 << A.log2n
    seeds, seed_off, ch, ch_off = build_inputs(A, rank, torch, np)
    d_seeds = (torch.from_numpy(seeds).to(dev), torch.from_numpy(seed_off).to(dev))
    d_ch = (torch.from_numpy(ch).to(dev).view(-1), torch.from_numpy(ch_off).to(dev))
    _, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, d_seeds, want_sk_coef=False, want_vk_coef=False, device=True)
    sig = eng.lm_sign(sch, sk_ntt, d_ch, device=True)
    del sk_ntt
    torch.cuda.synchronize()
    bad = torch.arange(0, n, 64, device=dev)
    sig.view(torch.int16)[bad, bad % p['l'], (3 * bad) % D] += 1
    torch.cuda.synchronize()
    verdict = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = _timed(torch, dist, world, dev, lambda: eng.lm_verify(sch, vk_ntt, d_ch, sig, p['vf_bd'], p['vf_wt'], out=verdict))
    expect = torch.ones(n, dtype=torch.uint8, device=dev)
    expect[bad] = 0
    ok = bool(torch.equal(verdict, expect))
    eng.profile(True)
    eng.profile_reset()
    eng.lm_verify(sch, vk_ntt, d_ch, sig, p['vf_bd'], p['vf_wt'], out=verdict)
    v_ms, _ = eng.profile_read('verify')
    s_ms, _ = eng.profile_read('sampler')
    eng.profile(False)
    eng.close()
    return {'verifies_per_s': world * n / (ms * 1e-3), 'ms': ms, 'n_per_gpu': n, 'verdicts_as_constructed': ok,
            'k_verify_ms': v_ms, 'k_sampler_ms': s_ms,
            'keccak_gperm_s_per_gpu': n * 26 / (s_ms * 1e-3) / 1e9, 'keccak_roofline_gperm_s': 4.28}


def secondary_bklm(a, rank, local, world, torch, np, dist):
    """configs[3]: aggregate + aggregate-verify of 2^k signatures per aggregate (secpar 128), the sorted list
    sharded over the ranks, ONE reduce of int32 partial sums each."""
    from lattice_cryptography_b200.distributed import reduce_partial, shard_range
    class A: secpar, log2n = 128, a.bklm_log2n
    eng, sch, p = _engine(128, local, np)
    dev = f'cuda:{local}'
    total = 1 << A.log2n
    start, count = shard_range(total, rank, world)
    # synthetic keys / 32-bit messages; object addresses ascend, so sorted order == index order
    rng = np.random.default_rng(777)
    bits = rng.integers(0, 2, (total, 32), dtype=np.uint8) + ord('0')
    msgs = [bytes(r).decode() for r in bits]
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * i:010x}>' for i in range(total)]
    agmsg = np.frombuffer(('[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']').encode(), dtype=np.uint8)
    d_agmsg = torch.from_numpy(agmsg.copy()).to(dev)
    seeds = [bin((0x9E3779B97F4A7C15 * (i + 1)) % (1 << 128))[2:].zfill(128) for i in range(start, start + count)]
    chm = [k + ', ' + m for k, m in zip(ident[start:start + count], msgs[start:start + count])]
    from lattice_cryptography_b200 import ragged
    cb, co = ragged(chm)
    d_chm = (torch.from_numpy(cb.copy()).to(dev), torch.from_numpy(co).to(dev))
    _, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, seeds, want_sk_coef=False, want_vk_coef=False, device=True)
    sigs = eng.lm_sign(sch, sk_ntt, d_chm, device=True)
    res = {}

    def agg():
        ag = eng.agg_coefs(sch, d_agmsg, start, count, device=True)
        part = eng.aggregate_partial(sch, sigs, ag, device=True)
        part = reduce_partial(part)
        if rank == 0:
            res['ag_sig'] = eng.aggregate_finish(part, device=True)
    ms_agg = _timed(torch, dist, world, dev, agg, reps=1)
    ag_sig = res.get('ag_sig')
    if world > 1:
        if rank != 0:
            ag_sig = torch.empty((p['l'], D), dtype=torch.int16, device=dev)
        dist.broadcast(ag_sig.view(torch.uint8), src=0)      # NCCL has no int16
    avf_bd = min(p['q'] // 2, total * p['vf_bd'])

    def aggv():
        ag = eng.agg_coefs(sch, d_agmsg, start, count, device=True)
        part = eng.aggverify_partial(sch, vk_ntt, d_chm, ag, device=True)
        part = reduce_partial(part)
        if rank == 0:
            res['ok'] = eng.aggverify_finish(part, ag_sig, total, total, avf_bd, 256)
    ms_aggv = _timed(torch, dist, world, dev, aggv, reps=1)
    eng.profile(True)
    eng.profile_reset()
    eng.agg_coefs(sch, d_agmsg, start, count, device=True)
    coef_ms, _ = eng.profile_read('agg_coefs')
    eng.profile(False)
    eng.close()
    perms = count * ((len(agmsg) + 12) // 136 + 1)
    return {'sigs_per_aggregate': total, 'aggregate_sigs_per_s': total / (ms_agg * 1e-3),
            'aggregate_verify_sigs_per_s': total / (ms_aggv * 1e-3), 'aggregate_ms': ms_agg,
            'aggregate_verify_ms': ms_aggv, 'verdict': res.get('ok') if rank == 0 else None,
            'agmsg_bytes': int(len(agmsg)), 'agg_coefs_ms_rank0': coef_ms,
            'agg_coefs_gperm_s_per_gpu': perms / (coef_ms * 1e-3) / 1e9, 'keccak_roofline_gperm_s': 4.28}


def secondary_single_ops(a, rank, local, world, torch, np, dist):
    """configs[0]: one key / one signature / one verification through the drop-in Python API (the reference's own
    calls, benchmarks/demo_signing.py), median wall-clock latency of 20 calls per secpar on rank 0."""
    if rank != 0:
        return None
    import statistics
    from lattice_cryptography_b200 import lattice_algebra as gla
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    gla.set_default_device(local)
    out = {}
    for secpar in (128, 256):
        pp = lm.make_setup_parameters(secpar)
        key = lm.keygen(pp=pp, num_keys_to_gen=1)[0]
        sig = lm.sign(pp=pp, otk=key, msg='QRL is awesome!')
        assert lm.verify(pp=pp, otvk=key[2], msg='QRL is awesome!', sig=sig)
        t = {'keygen': [], 'sign': [], 'verify': []}
        for _ in range(20):
            t0 = time.perf_counter()
            key = lm.keygen(pp=pp, num_keys_to_gen=1)[0]
            t1 = time.perf_counter()
            sig = lm.sign(pp=pp, otk=key, msg='QRL is awesome!')
            t2 = time.perf_counter()
            ok = lm.verify(pp=pp, otvk=key[2], msg='QRL is awesome!', sig=sig)
            t3 = time.perf_counter()
            assert ok
            t['keygen'].append(t1 - t0)
            t['sign'].append(t2 - t1)
            t['verify'].append(t3 - t2)
        out[f'secpar{secpar}_ms'] = {k: 1e3 * statistics.median(v) for k, v in t.items()}
    out['note'] = ('host-visible latency of single drop-in calls (Python objects in, Python objects out); the published '
                   'CPU log of the reference has 122 / 20 / 65-106 ms (secpar 128) for the same three calls')
    return out


def secondary_adaptor(a, rank, local, world, torch, np, dist):
    """configs[4]: adaptor pre-sign / pre-verify / adapt / verify / extract / witness-verify, 2^15 instances per GPU."""
    class A: secpar, log2n = 128, 15
    eng, sch, p = _engine(128, local, np)
    dev = f'cuda:{local}'
    n = 1 << A.log2n
    seeds, seed_off, ch, ch_off = build_inputs(A, rank, torch, np)
    d_seeds = (torch.from_numpy(seeds).to(dev), torch.from_numpy(seed_off).to(dev))
    d_ch = (torch.from_numpy(ch).to(dev).view(-1), torch.from_numpy(ch_off).to(dev))
    _, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, d_seeds, want_sk_coef=False, want_vk_coef=False, device=True)
    pvf_bd, vf_bd, ext_bd = p['vf_bd'], p['vf_bd'] + 1, 2 * p['vf_bd'] + 1
    st = {}
    t = {}
    t['witgen'] = _timed(torch, dist, world, dev, lambda: st.update(zip(('wit', 'st_ntt', 'st_coef'), eng.witgen(sch, d_seeds, device=True))))
    t['presign'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('presig', eng.lm_sign(sch, sk_ntt, d_ch, device=True)))
    t['preverify'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('pv', eng.lm_verify(sch, vk_ntt, d_ch, st['presig'], pvf_bd, 256, device=True)))
    t['adapt'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('sig', eng.vec_add(st['presig'], st['wit'], device=True)))
    t['verify'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('vv', eng.lm_verify(sch, vk_ntt, d_ch, st['sig'], vf_bd, 256, st_ntt=st['st_ntt'], device=True)))
    t['extract'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('ext', eng.vec_sub(st['sig'], st['presig'], device=True)))
    t['witness_verify'] = _timed(torch, dist, world, dev, lambda: st.__setitem__('wv', eng.witness_verify(st['ext'], st['st_ntt'], ext_bd, 256, device=True)))
    ok = bool(st['pv'].all().item() and st['vv'].all().item() and st['wv'].all().item())
    eng.close()
    return {'n_per_gpu': n, 'all_verdicts_true': ok,
            'ops_per_s': {k: world * n / (v * 1e-3) for k, v in t.items()}, 'ms': t}


def engine_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    p = SHIPPED[a.secpar]

    cpu = None
    if rank == 0 and a.gpus == 1 and not a.no_cpu_baseline:
        cpu = cpu_leg(a.secpar, a.cpu_per_core)       # before CUDA is initialised (forks workers)

    import numpy as np
    import torch
    import torch.distributed as dist
    from lattice_cryptography_b200 import Engine, make_scheme

    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    eng = Engine(a.secpar, p['q'], D, p['l'], device=local)
    eng.use_torch_stream()
    sch = make_scheme(sk_bd=p['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=p['ch_wt'])
    key_ch, _ = eng.hash2polyvec('KEY_CH_SEED', [KEY_CH_SEED], p['q'] // 2, D, p['l'])
    eng.set_key_ch(np.ascontiguousarray(key_ch[0]))

    n = 1 << a.log2n
    seeds, seed_off, ch, ch_off = build_inputs(a, rank, torch, np)
    dev = f'cuda:{local}'
    d_ch = torch.from_numpy(ch).to(dev)
    d_ch_off = torch.from_numpy(ch_off).to(dev)
    # keys and signatures come from the engine itself (tests/ pin them to the oracle)
    t0 = time.perf_counter()
    _, sk_ntt, vk_ntt, _ = eng.lm_keygen(sch, (torch.from_numpy(seeds).to(dev), torch.from_numpy(seed_off).to(dev)),
                                         want_sk_coef=False, want_vk_coef=False, device=True)
    torch.cuda.synchronize()
    t_keygen = time.perf_counter() - t0
    t0 = time.perf_counter()
    sig = eng.lm_sign(sch, sk_ntt, (d_ch.view(-1), d_ch_off), device=True)
    torch.cuda.synchronize()
    t_sign = time.perf_counter() - t0
    del sk_ntt
    torch.cuda.empty_cache()
    # every 64th triple is tampered: coefficient +1 / coefficient out of bound / message byte flipped
    bad = torch.arange(0, n, 64, device=dev)
    kind = (bad // 64) % 3
    sig_v = sig.view(torch.int16)
    i0, i1 = bad[kind == 0], bad[kind == 1]
    sig_v[i0, i0 % p['l'], (3 * i0) % D] += 1
    sig_v[i1, (i1 + 1) % p['l'], (7 * i1) % D] = p['vf_bd'] + 1
    i2 = bad[kind == 2]
    d_ch[i2, -1] ^= 1           # '0' <-> '1'
    expect = torch.ones(n, dtype=torch.uint8, device=dev)
    expect[bad] = 0
    verdict = torch.empty(n, dtype=torch.uint8, device=dev)
    chm = (d_ch.view(-1), d_ch_off)

    def step():
        eng.lm_verify(sch, vk_ntt, chm, sig, p['vf_bd'], p['vf_wt'], out=verdict)

    step()
    torch.cuda.synchronize()
    if not torch.equal(verdict, expect):
        raise SystemExit(f'rank {rank}: verdicts differ from the construction rule '
                         f'({int((verdict != expect).sum())} of {n})')

    for _ in range(a.warmup):
        step()
    eng.profile(True)
    eng.profile_reset()
    launches0 = eng.launch_count
    clocks = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if clocks else None
    launches = eng.launch_count - launches0
    v_ms, v_n = eng.profile_read('verify')
    s_ms, s_n = eng.profile_read('sampler')
    eng.profile(False)
    total_ms = float(ms.item())
    value = world * n * a.steps / (total_ms * 1e-3)

    # ---- end to end: the same call with pinned HOST buffers (H2D of every input + D2H of the verdicts inside)
    h_sig = torch.empty(sig.shape, dtype=torch.int16, pin_memory=True)
    h_sig.copy_(sig_v)
    h_vk = torch.empty(vk_ntt.shape, dtype=torch.uint16, pin_memory=True)
    h_vk.copy_(vk_ntt)
    h_ch = torch.empty(d_ch.shape, dtype=torch.uint8, pin_memory=True)
    h_ch.copy_(d_ch)
    h_off = torch.empty(d_ch_off.shape, dtype=torch.int64, pin_memory=True)
    h_off.copy_(d_ch_off)
    h_verdict = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    np_sig, np_vk, np_ch, np_off, np_verdict = (h_sig.numpy(), h_vk.view(torch.int16).numpy().view(np.uint16),
                                                h_ch.numpy().reshape(-1), h_off.numpy(), h_verdict.numpy())

    def e2e_step():
        eng.lm_verify(sch, np_vk, (np_ch, np_off), np_sig, p['vf_bd'], p['vf_wt'], out=np_verdict)

    e2e_step()
    if not np.array_equal(np_verdict, expect.cpu().numpy()):
        raise SystemExit(f'rank {rank}: end-to-end verdicts differ from the construction rule')
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n * a.e2e_steps / float(e2e_s.item())
    h2d = np_sig.nbytes + np_vk.nbytes + np_ch.nbytes + np_off.nbytes
    d2h = np_verdict.nbytes

    # ---- the same end-to-end call on the packed wire format (opt-in): 11/13-bit signature
    # coefficients and 14/16-bit key slots cross PCIe instead of 16-bit ones
    import math
    sbits, kbits = math.ceil(math.log2(2 * p['vf_bd'] + 1)), math.ceil(math.log2(p['q']))
    d_sig_p = eng.pack(sig_v, sbits, p['vf_bd'], device=True)
    d_vk_p = eng.pack(vk_ntt, kbits, 0, device=True)
    eng.lm_verify_packed(sch, d_vk_p, kbits, chm, d_sig_p, sbits, p['vf_bd'], p['vf_bd'], p['vf_wt'], out=verdict)
    torch.cuda.synchronize()
    if not torch.equal(verdict, expect):
        raise SystemExit(f'rank {rank}: packed verdicts differ from the construction rule')
    pk0, pk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pk0.record()
    eng.lm_verify_packed(sch, d_vk_p, kbits, chm, d_sig_p, sbits, p['vf_bd'], p['vf_bd'], p['vf_wt'], out=verdict)
    pk1.record()
    torch.cuda.synchronize()
    packed_resident_ms = pk0.elapsed_time(pk1)
    h_sig_p = torch.empty(d_sig_p.shape, dtype=torch.uint8, pin_memory=True)
    h_sig_p.copy_(d_sig_p)
    h_vk_p = torch.empty(d_vk_p.shape, dtype=torch.uint8, pin_memory=True)
    h_vk_p.copy_(d_vk_p)
    torch.cuda.synchronize()
    del d_sig_p, d_vk_p
    np_sig_p, np_vk_p = h_sig_p.numpy(), h_vk_p.numpy()

    def e2e_packed_step():
        eng.lm_verify_packed(sch, np_vk_p, kbits, (np_ch, np_off), np_sig_p, sbits, p['vf_bd'], p['vf_bd'], p['vf_wt'],
                             out=np_verdict)

    np_verdict[:] = 2
    e2e_packed_step()
    if not np.array_equal(np_verdict, expect.cpu().numpy()):
        raise SystemExit(f'rank {rank}: packed end-to-end verdicts differ from the construction rule')
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        e2e_packed_step()
    e2e_p_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_p_s, op=dist.ReduceOp.MAX)
    e2e_packed = {'value': world * n * a.e2e_steps / float(e2e_p_s.item()), 'unit': UNIT,
                  'h2d_bytes_per_step': np_sig_p.nbytes + np_vk_p.nbytes + np_ch.nbytes + np_off.nbytes,
                  'd2h_bytes_per_step': d2h, 'steps': a.e2e_steps, 'sig_bits': sbits, 'key_bits': kbits,
                  'resident_ms_per_step': packed_resident_ms,
                  'note': 'lcb_lm_verify_packed_batch on pinned host buffers; opt-in wire format, not the headline e2e'}
    del h_sig_p, h_vk_p, np_sig_p, np_vk_p

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = peaks.get('hbm_gbs', 6650.0)
        # algorithmic bytes of one verify inside k_verify: signature + vk + challenge pairs + verdict
        unit_bytes = p['l'] * D * 2 + 2 * D * 2 + p['ch_wt'] * 4 + 1
        k_ms = v_ms / max(v_n, 1)
        achieved = unit_bytes * n / (k_ms * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': total_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u32 (mod-q integer arithmetic, q < 2^16)', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'secpar': a.secpar, 'q': p['q'], 'd': D, 'l': p['l'],
                       'triples_per_gpu': n, 'tampered_every': 64,
                       'l2': f'inputs {(sig.numel() * 2 + vk_ntt.numel() * 2 + d_ch.numel()) / 1e9:.2f} GB per pass '
                             f'>> 126 MB L2, no flush needed'},
            'roofline': {'bound': 'hbm', 'kernel': 'k_verify', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak,
                         # DRAM bytes per launch from the ncu --set full capture of this kernel
                         # (profiles/prof_r1_verify.summary.txt: 2.0344 GB read + 5.0 MB written for 2^18
                         # verifies = 7,780 B per verify), scaled to this launch's batch
                         'traffic': 7780 * n if a.secpar == 128 else None,
                         'algorithmic_bytes_per_launch': unit_bytes * n,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback',
                         'algorithmic_bytes_per_unit': unit_bytes,
                         'kernel_ms_per_launch': k_ms, 'kernel_share_of_step': v_ms / total_ms,
                         'sampler_ms_per_launch': s_ms / max(s_n, 1),
                         'note': 'k_verify is integer-issue bound by design (see DESIGN.md); HBM fraction is the '
                                 'contract figure, int-pipe figures are in int_pipe'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': a.e2e_steps, 'host_cpus_bound_per_rank': numa_cpus},
            'e2e_packed': e2e_packed,
            'gpu_launches': launches,
            'clocks': clk,
            'setup': {'keygen_s': t_keygen, 'sign_s': t_sign, 'keygen_keys_per_s': n / t_keygen,
                      'sign_sigs_per_s': n / t_sign,
                      'note': 'cold first calls of this process (memory-pool growth and output allocation inside); '
                              'warm rates are in secondary and profiles/README.md'},
        }
        # instruction-issue view of the same kernel: warp-instructions per verify from the ncu count of this
        # build (profiles/prof_r1_verify.summary.txt: 1,004.7 M for 2^18 verifies), against the issue limit of
        # 4 schedulers x 148 SMs x 1 warp-instruction per clock at the max clock
        instr_per_unit = 32 * 3833 if a.secpar == 128 else None
        clk_hz = ((clk or {}).get('sm_max_mhz') or 1965.0) * 1e6
        if instr_per_unit:
            achieved_t = instr_per_unit * n / (k_ms * 1e-3) / 1e12
            issue_peak = 148 * 128 * clk_hz / 1e12
            line['roofline']['int_pipe'] = {
                'thread_instr_per_unit': instr_per_unit, 'achieved_tinstr_s': achieved_t,
                'issue_peak_tinstr_s': issue_peak, 'issue_frac': achieved_t / issue_peak,
                'ncu_pipe_utilisation': {'issue_active': 0.751, 'alu': 0.523, 'fma': 0.355, 'lsu': 0.342},
                'note': 'k_verify is bound by instruction issue: 5-instruction FP32-assisted butterflies spread '
                        'over the ALU, FMA-heavy and FMA-lite pipes (DESIGN.md 3.3)'}
        if cpu is not None:
            line['cpu_baseline'] = cpu
    del sig, sig_v, vk_ntt, d_ch, h_sig, h_vk, h_ch, np_sig, np_vk, np_ch
    torch.cuda.empty_cache()
    secondary = {}
    if not a.no_secondary:
        for name, fn in (('keygen_sign_secpar256', secondary_keygen_sign),
                         ('lm_verify_secpar256', secondary_verify_secpar256), ('bklm', secondary_bklm),
                         ('adaptor', secondary_adaptor), ('single_ops', secondary_single_ops)):
            try:
                secondary[name] = fn(a, rank, local, world, torch, np, dist)
            except Exception as exc:          # the headline line must survive a secondary failure
                secondary[name] = {'error': repr(exc)[:300]}
    if rank == 0:
        line['secondary'] = secondary
        bk = secondary.get('bklm') or {}
        if 'aggregate_verify_sigs_per_s' in bk:       # the second half of BASELINE.json's metric, on the same line
            line['also'] = {'metric': 'BKLM aggregate-verify sigs/sec', 'value': bk['aggregate_verify_sigs_per_s'],
                            'unit': 'signatures/s', 'n_gpus': world,
                            'config': {'workload': f"bklm aggregate-verify, N = {bk['sigs_per_aggregate']} signatures per "
                                                   f"aggregate sharded over {world} GPU(s), secpar 128, exact reference "
                                                   f"semantics incl. the O(N^2) aggregation-coefficient hashing"},
                            'aggregate_sigs_per_s': bk.get('aggregate_sigs_per_s')}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        engine_arm(args)
