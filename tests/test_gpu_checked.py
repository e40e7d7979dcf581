"""The checked build (make CHECKED=1 -> liblcb200_checked.so, -DLCB_CHECKED) and guard bands around outputs: the
stand-in for compute-sanitizer, which is closed on the GPU pool (profiles/sanitizer_r2.txt).

* every hand-rolled bound of a global access carries a device-side assert in that build (the aligned-word windows of
  InputView::load_block and of both k_agg_coefs kernels, the staged rows of k_verify, the decoder's position select);
  the edge-case tests (rate-block boundaries, alignments, ragged lengths, digit-count changes, packed rows, generic
  degrees) are re-run on it in a subprocess - a failed check traps the kernel and fails the test;
* a self-test proves the asserts are live in that library and compiled out of the production one;
* output buffers are placed between canary bands, which must survive every call.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, 'lattice_cryptography_b200', 'liblcb200_checked.so')


def _checked_lib():
    if not os.path.exists(CHECKED):
        subprocess.run(['make', '-C', os.path.join(ROOT, 'lattice_cryptography_b200', 'csrc'), '-j4', 'CHECKED=1'], check=True,
                       capture_output=True)
    return CHECKED


def _run(code, lib=None):
    env = dict(os.environ)
    if lib:
        env['LCB200_LIB'] = lib
    return subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)


SELFTEST = ('import sys; sys.path.insert(0, %r)\n'
            'from lattice_cryptography_b200 import Engine, _ffi\n'
            'lib = _ffi.load(); e = Engine(128, 11777, 256, 13)\n'
            'print("checked", lib.lcb_build_is_checked(), "selftest", lib.lcb_checked_selftest(e._ctx))\n' % ROOT)


def test_asserts_are_live_in_the_checked_build_only():
    r = _run(SELFTEST)
    assert 'checked 0 selftest 0' in r.stdout, r.stdout + r.stderr
    r = _run(SELFTEST, _checked_lib())
    assert 'checked 1 selftest -2' in r.stdout, r.stdout + r.stderr          # LCB_ERR_CUDA: device-side assert


def test_edge_cases_pass_on_the_checked_build():
    sel = ('boundaries or agg_coefs or field_widths or packed or golden or seed_lengths or sampler or '
           'keygen_sign_verify or bklm_and_adaptor or non_monomial or device_buffers')
    env = dict(os.environ, LCB200_LIB=_checked_lib())
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', 'tests/test_gpu_wire.py',
                        'tests/test_gpu_generic.py', 'tests/test_gpu_wide.py', '-q', '-x', '-m', 'gpu', '-k', sel,
                        '-p', 'no:cacheprovider'], env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = r.stdout[-2500:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and 'failed' not in r.stdout, tail


class _Banded(object):
    """Engine whose outputs live between two 4 KiB canary bands (device memory)."""
    PAD = 4096

    def __init__(self, eng):
        import torch
        self.eng, self.torch, self.bands = eng, torch, []
        self._orig = eng._out

        def banded(shape, dtype, device):
            n = int(np.prod(shape)) if len(shape) else 1
            tdt = {np.int16: torch.int16, np.uint16: torch.uint16, np.uint8: torch.uint8, np.int32: torch.int32,
                   np.uint32: torch.uint32, np.int64: torch.int64}[dtype]
            item = torch.empty((), dtype=tdt).element_size()
            pad = self.PAD // item
            raw = torch.full(((2 * pad + n) * item,), 0xA5, dtype=torch.uint8, device=f'cuda:{eng.device}')
            view = raw[pad * item:(pad + n) * item].view(tdt).view(shape)
            self.bands.append((raw, pad * item, n * item))
            return view
        eng._out = banded

    def intact(self):
        for raw, lo, n in self.bands:
            head, tail = raw[:lo], raw[lo + n:]
            if not bool((head == 0xA5).all().item() and (tail == 0xA5).all().item()):
                return False
        return True

    def restore(self):
        self.eng._out = self._orig


@pytest.mark.parametrize('secpar,q,d,l', [(128, 11777, 256, 13), (256, 39937, 256, 23), (128, 12289, 512, 2), (128, 193, 32, 3)])
def test_outputs_stay_inside_their_buffers(secpar, q, d, l):
    from lattice_cryptography_b200 import Engine, make_scheme
    eng = Engine(secpar, q, d, l)
    b = _Banded(eng)
    try:
        sk_bd, ch_wt = (45, 20) if d == 256 else (2, 5)
        sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt, wit_bd=1, wit_wt=min(20, d))
        key_ch, _ = eng.hash2polyvec('KEY_CH_SEED', ['bands'], q // 2, d, l, device=True)
        eng.set_key_ch(key_ch[0].cpu().numpy())
        for n in (1, 7, 33):
            seeds = [bin(5 + j)[2:].zfill(secpar) for j in range(n)]
            chm = [f'<k{j}>, ' + 'm' * (j % 9) for j in range(n)]
            sk_coef, sk_ntt, vk_ntt, vk_coef = eng.lm_keygen(sch, seeds, device=True)
            sig = eng.lm_sign(sch, sk_ntt, chm, device=True)
            vf_bd = min(q // 2, sk_bd * (1 + ch_wt))
            verdict = eng.lm_verify(sch, vk_ntt, chm, sig, vf_bd, d, device=True)
            assert bool(verdict.all().item())
            pairs = eng.challenge(sch, chm, device=True)
            wit, st_ntt, st_coef = eng.witgen(sch, seeds, device=True)
            full = eng.vec_add(sig, wit, device=True)
            eng.vec_sub(full, sig, device=True)
            eng.witness_verify(wit, st_ntt, 1, d, device=True)
            ag = eng.agg_coefs(sch, b'[' + b'x' * (100 * n) + b']', 3, n, device=True)
            part = eng.aggregate_partial(sch, sig, ag, device=True)
            eng.aggregate_finish(part, device=True)
            eng.aggverify_partial(sch, vk_ntt, chm, ag, device=True)
            eng.ntt_inv(eng.ntt_fwd(vk_coef, device=True), device=True)
            eng.poly_mul(vk_coef, vk_coef, device=True)
            eng.shake256(chm, 200 + n, device=True)
            if d == 256:
                p = eng.pack(sig, 11 if secpar == 128 else 13, vf_bd, device=True)
                eng.unpack(p, 11 if secpar == 128 else 13, vf_bd, device=True)
            eng.synchronize()
            assert b.intact(), f'a canary band was overwritten at n = {n}'
        assert len(b.bands) > 40
    finally:
        b.restore()
        eng.close()
