"""acquire.SIGNALS (and the oracle's own SCRIPTS table) against the constants hard-coded in
the reference scripts — regex-extracted from /root/reference when it is present, and always
against the snapshot of that extraction committed as tests/golden/script_constants.json."""
import json
import os
import re

import pytest

from gnsstools import acquire
from oracle import acq_oracle as orc
from oracle import ref_lift

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'script_constants.json')


def extract(script):
    src = open(os.path.join(ref_lift.REF, 'acquire-%s.py' % script)).read()
    body = re.search(r'^def search\(.*?^  return [^\n]*\n', src, re.S | re.M).group(0)
    fs = eval(re.search(r'^\s+fs = (.+)$', body, re.M).group(1))
    n = eval(re.search(r'^\s+n = ([0-9*]+)', body, re.M).group(1))
    blocks = re.search(r'^\s+blocks = (.+)$', body, re.M)
    rng = re.search(r'for block in range\((.+?)\)', body).group(1)
    out = {
        'fs': fs, 'n': n,
        'blocks': blocks.group(1).strip() if blocks else rng,
        'pad': 'np.concatenate' in body,
        'boc': 'boc11' in body,
        'normalize': 'np.mean(q)' in body,
        'mod_L': re.search(r'm_code = m_code%', body) is not None,
        'carrier': (re.search(r'nco\.nco\(-\((\d+)\*chan', body) or [None, '0'])[1],
        'cutoff': float(re.search(r'firwin\(161,([0-9.e]+)/', src).group(1)),
        'prns': re.search(r'add_option\("--(?:prn|channel)", default="([^"]*)"', src).group(1),
        'doppler': re.search(r'add_option\("--doppler-search".*?default="([^"]*)"', src).group(1),
        'time': int(re.search(r'add_option\("--time".*?default=(\d+)', src).group(1)),
        'fmt': re.search(r"return '([^']*)' %", src).group(1),
        'per_ms': eval(re.search(r'np\.arange\(ms_pad\*([0-9*]+)\)', src).group(1)),
        'module': re.search(r'^import gnsstools\.(\w+\.\w+) as', src, re.M).group(1),
    }
    return out


def blocks_fn(expr):
    return lambda ms: eval(expr, {'ms': ms})


def check(name, c):
    s = acquire.SIGNALS[name]
    assert (s.fs, s.n, s.pad, s.boc, s.normalize, s.mod_L) == (c['fs'], c['n'], c['pad'], c['boc'], c['normalize'], c['mod_L']), name
    assert s.carrier_step == float(c['carrier']) and s.fdma == (c['carrier'] != '0')
    assert (s.cutoff, s.prns, s.doppler, s.time, s.fmt, s.module) == (c['cutoff'], c['prns'], c['doppler'], c['time'], c['fmt'], c['module']), name
    assert int(round(s.fs * 0.001)) == c['per_ms']
    o = orc.SCRIPTS[name]
    assert (o['fs'], o['n'], o['pad'], o['boc'], o['normalize'], o['mod_L'], o['carrier_step'], o['module']) == \
        (c['fs'], c['n'], c['pad'], c['boc'], c['normalize'], c['mod_L'], float(c['carrier']), c['module']), name
    for ms in (1, 4, 10, 20, 40, 80, 85):
        want = blocks_fn(c['blocks'])(ms)
        assert s.blocks(ms) == want and o['blocks'](ms) == want, (name, ms)


def test_tables_match_committed_snapshot():
    gold = json.load(open(GOLD))
    assert sorted(gold) == sorted(acquire.SIGNALS) == sorted(orc.SCRIPTS)
    for name, c in gold.items():
        check(name, c)


@pytest.mark.skipif(not ref_lift.available(), reason='/root/reference not present')
def test_snapshot_matches_reference_scripts():
    gold = json.load(open(GOLD))
    scripts = sorted(f[len('acquire-'):-3] for f in os.listdir(ref_lift.REF)
                     if f.startswith('acquire-') and f.endswith('.py'))
    serial = {'gps-l2cl', 'glonass-l1-p', 'glonass-l2-p'}        # non-FFT, out of scope
    assert sorted(set(scripts) - serial) == sorted(gold)
    for name in gold:
        assert extract(name) == gold[name], name


if __name__ == '__main__':       # regenerate the snapshot (build container only)
    names = sorted(f[len('acquire-'):-3] for f in os.listdir(ref_lift.REF) if f.startswith('acquire-') and f.endswith('.py'))
    snap = {n: extract(n) for n in names if n not in ('gps-l2cl', 'glonass-l1-p', 'glonass-l2-p')}
    json.dump(snap, open(GOLD, 'w'), indent=1, sort_keys=True)
    print('wrote', GOLD, len(snap))
