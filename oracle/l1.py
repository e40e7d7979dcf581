"""CPU ORACLE (test infrastructure): which `lattice_algebra` (the reference's L1 dependency) is in use, and a
U-item by U-item comparison of two of them.

The reference pins `lattice_algebra==0.1.1` (requirements.txt:1, setup.py:44) and that package is absent from
/root/reference and from the build image, so oracle/lattice_algebra/ RESTATES it (PARITY UNPINNED).
The moment a real copy is reachable it takes precedence everywhere the oracle is built from:

  search order (find_real):  $LCB_LATTICE_ALGEBRA (a directory that contains the package directory
                             `lattice_algebra/`), <repo>/baseline/_ref, then any `lattice_algebra` importable
                             from site-packages with oracle/ taken off sys.path.

* oracle/ref_loader.py puts the real package ahead of the restatement when it loads the reference's modules;
* oracle/gen_golden.py records which L1 produced tests/golden/ in golden.json["l1"]
  ("restated" | "lattice_algebra==<version>");
* tests/test_oracle_pin.py runs `probe()` under both and FAILS, naming the U-items, if they disagree.

`probe()` is executed in a subprocess per L1 (two packages of the same name cannot share one interpreter) and
only uses the ten names the reference itself imports, plus a few internals when they exist.

    python oracle/l1.py which
    python oracle/l1.py probe [<dir containing lattice_algebra/> | restated]      # JSON on stdout
    python oracle/l1.py compare                                                    # exit 1 on any disagreement
"""
import importlib
import importlib.util
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RESTATED = 'restated'


def _is_pkg_dir(d: str) -> bool:
    return bool(d) and (os.path.isfile(os.path.join(d, 'lattice_algebra', '__init__.py')) or
                        os.path.isfile(os.path.join(d, 'lattice_algebra.py')))


def find_real():
    """Directory to put on sys.path so that `import lattice_algebra` yields the REAL package, or None."""
    env = os.environ.get('LCB_LATTICE_ALGEBRA')
    if env:
        if not _is_pkg_dir(env):
            raise RuntimeError(f'LCB_LATTICE_ALGEBRA={env} does not contain a lattice_algebra package')
        return os.path.abspath(env)
    cand = os.path.join(ROOT, 'baseline', '_ref')
    if _is_pkg_dir(cand):
        return cand
    for p in sys.path:
        if not p or os.path.abspath(p) == HERE:
            continue
        if _is_pkg_dir(p):
            return os.path.abspath(p)
    return None


def label_of(path) -> str:
    """'restated' or 'lattice_algebra==<version>' (version from the package's metadata if it has any)."""
    if path in (None, RESTATED):
        return RESTATED
    version = 'unknown'
    for name in os.listdir(path):
        low = name.lower()
        if low[:15] in ('lattice_algebra', 'lattice-algebra') and low[15:16] == '-' and \
                (low.endswith('.dist-info') or low.endswith('.egg-info')):
            version = low[16:].rsplit('.', 1)[0].split('-')[0]
    return f'lattice_algebra=={version}'


def current_label() -> str:
    return label_of(find_real())


# ------------------------------------------------------------------------------------------------ the probe
U_ITEMS = {
    'U7_lattice_parameters': 'LatticeParameters: rou / rou_inv / halfmod / logmod / n',
    'U1_U6_hash2polynomialvector': 'salt order, digest-to-bits order, chunking, index and coefficient decoding',
    'U1_U6_hash2polynomial': 'the same through hash2polynomial (challenge and aggregation-coefficient shapes)',
    'U8_U9_U11_operators': 'get_coef_rep, ==, +, -, *, vector dot product, vector ** polynomial, sum()',
    'U13_constructor': 'Polynomial(lp, coefs, const_time_flag) and PolynomialVector(lp, entries, ...)',
    'L3_ntt_representation': 'raw 2d-point ntt_representation list of a polynomial',
    'U12_predicates': 'is_bitstring, is_ntt_friendly_prime, UNIFORM_INFINITY_WEIGHT',
    'U2_binary_digest': 'binary_digest(msg, num_bytes, salt) when the package exposes it',
}


def _rep(x):
    """JSON-able form of a get_coef_rep() result (dict keys become strings, tuples lists)."""
    if isinstance(x, dict):
        return {str(k): _rep(v) for k, v in sorted(x.items())}
    if isinstance(x, (list, tuple)):
        return [_rep(v) for v in x]
    return x


def probe() -> dict:
    """Run every probe against whatever `import lattice_algebra` resolves to in THIS interpreter."""
    la = importlib.import_module('lattice_algebra')
    out = {'_file': getattr(la, '__file__', '?')}

    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:                     # a probe that cannot run is reported, not hidden
            out[name] = {'error': f'{type(exc).__name__}: {exc}'[:300]}

    grids = [(11777, 256, 13), (39937, 256, 23), (193, 32, 2)]
    lps = {}

    def params():
        rows = []
        for q, d, l in grids:
            lp = la.LatticeParameters(modulus=q, degree=d, length=l)
            lps[(q, d, l)] = lp
            rows.append([getattr(lp, a, None) for a in ('modulus', 'degree', 'length', 'rou', 'rou_inv', 'halfmod', 'logmod', 'n')])
        return rows
    guard('U7_lattice_parameters', params)

    dist = la.UNIFORM_INFINITY_WEIGHT

    def bti(secpar, d, wt):
        from math import ceil, log2
        return ceil(log2(d)) + (wt - 1) * (ceil(log2(d)) + secpar)

    def btd(secpar, bd):
        from math import ceil, log2
        return ceil(log2(bd)) + 1 + secpar

    def h2pv():
        rows = []
        for (q, d, l), secpar, bd, wt, salt, msg in [((11777, 256, 13), 128, 45, 256, 'SK_SALTLEFT', '0' * 127 + '1'),
                                                    ((39937, 256, 23), 256, 65, 256, 'SK_SALTRIGHT', '01' * 128),
                                                    ((11777, 256, 13), 128, 1, 20, 'WIT_SALT', '1' * 128),
                                                    ((193, 32, 2), 128, 3, 7, 'S', 'abc'),
                                                    ((11777, 256, 13), 128, 5888, 256, 'KEY_CH_SEED', 'lcb200 golden key_ch v1')]:
            v = la.hash2polynomialvector(secpar=secpar, lp=lps[(q, d, l)], distribution=dist, dist_pars={'bd': bd, 'wt': wt},
                                         num_coefs=wt, bti=bti(secpar, d, wt), btd=btd(secpar, bd), msg=msg, salt=salt,
                                         const_time_flag=False)
            rows.append(_rep(v.get_coef_rep()))
        return rows
    guard('U1_U6_hash2polynomialvector', h2pv)

    polys = []

    def h2p():
        rows = []
        for (q, d, l), secpar, bd, wt, salt, msg in [((11777, 256, 13), 128, 1, 20, 'CH_SALT', '<key object at 0x7f0000000010>, QRL is awesome!'),
                                                    ((39937, 256, 23), 256, 1, 50, 'CH_SALT', 'x' * 349),
                                                    ((11777, 256, 13), 128, 1, 1, 'AG_SALT7', "[(<k>, '0101')]"),
                                                    ((11777, 256, 13), 128, 1, 1, 'AG_SALT12345', 'y' * 1000),
                                                    ((11777, 256, 13), 128, 45, 256, 'P', ''),
                                                    ((193, 32, 2), 128, 96, 32, 'Q', 'm')]:
            p = la.hash2polynomial(secpar=secpar, lp=lps[(q, d, l)], distribution=dist, dist_pars={'bd': bd, 'wt': wt},
                                   salt=salt, msg=msg, num_coefs=wt, bti=bti(secpar, d, wt), btd=btd(secpar, bd),
                                   const_time_flag=False)
            polys.append(((q, d, l), p))
            rows.append(_rep(p.get_coef_rep()))
        return rows
    guard('U1_U6_hash2polynomial', h2p)

    def ops():
        lp = lps[(11777, 256, 13)]
        a, b = polys[0][1], polys[4][1]
        mk = lambda s: la.hash2polynomialvector(secpar=128, lp=lp, distribution=dist, dist_pars={'bd': 45, 'wt': 256},
                                                num_coefs=256, bti=bti(128, 256, 256), btd=btd(128, 45), msg=s,
                                                salt='V', const_time_flag=False)
        u, v = mk('u'), mk('v')
        rows = {'a*b': _rep((a * b).get_coef_rep()), 'a+b': _rep((a + b).get_coef_rep()),
                'a-b': _rep((a - b).get_coef_rep()), 'u*v': _rep((u * v).get_coef_rep()),
                'u**a_first': _rep((u ** a).get_coef_rep()[0]), 'u+v_last': _rep((u + v).get_coef_rep()[-1]),
                'u-v_first': _rep((u - v).get_coef_rep()[0]),
                'sum': _rep(sum([a, b, a]).get_coef_rep()),
                'eq': [a == a, a == b, u == u, u == v, (a * b) == (b * a), (u ** a) * v == (u * v) * a],
                'vec_entries': len(u.entries), 'lin': (u * (v ** a + u)) == ((u * v) * a + u * u)}
        small = lps[(193, 32, 2)]
        c = polys[5][1]
        rows['small_c*c'] = _rep((c * c).get_coef_rep())
        del small
        return rows
    guard('U8_U9_U11_operators', ops)

    def ctor():
        lp = lps[(193, 32, 2)]
        p = la.Polynomial(lp, {0: 1, 5: -2, 31: 96}, False)
        w = la.PolynomialVector(lp, [p, p * p], False)
        return [_rep(p.get_coef_rep()), _rep(w.get_coef_rep()), _rep((w * w).get_coef_rep())]
    guard('U13_constructor', ctor)

    guard('L3_ntt_representation', lambda: [[int(x) for x in polys[i][1].ntt_representation] for i in (0, 5)])
    guard('U12_predicates', lambda: [la.is_bitstring(''), la.is_bitstring('0110'), la.is_bitstring('012'),
                                     la.is_bitstring(5), la.is_ntt_friendly_prime(modulus=193, degree=32),
                                     la.is_ntt_friendly_prime(modulus=197, degree=32),
                                     la.is_ntt_friendly_prime(modulus=11777, degree=256), dist])
    if hasattr(la, 'binary_digest'):
        guard('U2_binary_digest', lambda: [la.binary_digest('abc', 5, 'salt'), la.binary_digest('', 1, ''),
                                           la.binary_digest('m' * 200, 3, 'SK_SALTLEFT')])
    return out


def run_probe(path) -> dict:
    """probe() in a fresh interpreter whose `lattice_algebra` is the package under `path` (or the restatement)."""
    first = HERE if path in (None, RESTATED) else path
    code = ('import sys, json; sys.path.insert(0, %r); sys.path.insert(1, %r); '
            'sys.modules.pop("lattice_algebra", None); import l1; print(json.dumps(l1.probe()))' % (first, HERE))
    if first != HERE:
        # the probe module itself lives in oracle/, which also holds the restatement: import l1 by file path instead
        code = ('import sys, json, importlib.util; sys.path.insert(0, %r); '
                'spec = importlib.util.spec_from_file_location("l1", %r); l1 = importlib.util.module_from_spec(spec); '
                'spec.loader.exec_module(l1); print(json.dumps(l1.probe()))' % (first, os.path.join(HERE, 'l1.py')))
    env = {k: v for k, v in os.environ.items() if k != 'PYTHONPATH'}
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd='/')
    if r.returncode != 0:
        raise RuntimeError(f'probe under {first} failed:\n{r.stderr[-2000:]}')
    return json.loads(r.stdout.strip().splitlines()[-1])


def compare(real_path, restated_path=RESTATED):
    """-> (disagreeing U-items, both probe results).  An item missing on one side (optional internals) is skipped;
    an item that raised on one side counts as a disagreement."""
    a, b = run_probe(real_path), run_probe(restated_path)
    bad = [k for k in U_ITEMS if k in a and k in b and a[k] != b[k]]
    return bad, a, b


if __name__ == '__main__':
    cmd = sys.argv[1] if len(sys.argv) > 1 else 'which'
    if cmd == 'which':
        p = find_real()
        print(json.dumps({'l1': label_of(p), 'path': p or os.path.join(HERE, 'lattice_algebra')}))
    elif cmd == 'probe':
        print(json.dumps(run_probe(sys.argv[2] if len(sys.argv) > 2 else RESTATED)))
    elif cmd == 'compare':
        real = find_real()
        if real is None:
            print('no real lattice_algebra reachable: parity stays UNPINNED (restated L1 only)')
            sys.exit(0)
        bad, a, b = compare(real)
        for k in bad:
            print(f'DISAGREE {k}: {U_ITEMS[k]}')
        print(f'{label_of(real)} at {real}: {len(U_ITEMS) - len(bad)} of {len(U_ITEMS)} items agree')
        sys.exit(1 if bad else 0)
