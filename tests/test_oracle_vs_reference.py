"""Pin the oracle restatement against the reference's own search() (build container only)."""
import zlib

import numpy as np
import pytest

from oracle import acq_oracle as orc
from oracle import ref_lift

pytestmark = pytest.mark.skipif(not ref_lift.available(), reason='/root/reference not present')

# (script, key(prn/chan), doppler grid, ms)
CASES = [
    ('gps-l1', 1, (-5000, 5000, 500), 1),
    ('gps-l1', 7, (-1000, 1000, 250), 3),
    ('xona-x1', 0, (-600, 600, 200), 2),
    ('xona-x5p', 0, (-400, 400, 200), 1),
    ('glonass-l1', -3, (-400, 400, 200), 2),
    ('glonass-l2', 5, (-400, 400, 200), 1),
    ('gps-l1cd', 3, (-40, 40, 20), 10),
    ('gps-l1cp', 4, (-40, 40, 20), 10),
    ('beidou-b1cd', 5, (-40, 40, 20), 10),
    ('beidou-b1cp', 6, (-40, 40, 20), 10),
    ('galileo-e1b', 11, (-100, 100, 50), 8),
    ('galileo-e1c', 12, (-100, 100, 50), 8),
    ('beidou-b1i', 8, (-400, 400, 200), 2),
    ('beidou-b2i', 9, (-400, 400, 200), 2),
    ('gps-l2cm', 5, (-40, 40, 20), 40),
    ('gps-l5i', 1, (-400, 400, 200), 2),
    ('gps-l5q', 2, (-400, 400, 200), 1),
    ('galileo-e5ai', 3, (-400, 400, 200), 1),
    ('galileo-e5aq', 4, (-400, 400, 200), 1),
    ('galileo-e5bi', 5, (-400, 400, 200), 1),
    ('galileo-e5bq', 6, (-400, 400, 200), 1),
    ('galileo-e6b', 7, (-400, 400, 200), 2),
    ('galileo-e6c', 8, (-400, 400, 200), 2),
    ('beidou-b2ap', 9, (-400, 400, 200), 1),
    ('beidou-b2bi', 20, (-400, 400, 200), 1),
    ('beidou-b2bq', 21, (-400, 400, 200), 1),
    ('beidou-b3i', 10, (-400, 400, 200), 1),
    ('glonass-l3ocd', 11, (-400, 400, 200), 1),
    ('glonass-l3ocp', 12, (-400, 400, 200), 1),
]


def ref_chips(ns, script, key):
    s = orc.SCRIPTS[script]
    mod = [v for v in ns.values() if getattr(v, '__name__', '') == 'gnsstools.' + s['module']][0]
    fn = getattr(mod, s['module'].split('.')[-1] + '_code')
    return np.asarray(fn() if s['fdma'] else fn(key), dtype=np.float64), mod


def synth(script, chips, key, ms, seed):
    s = orc.SCRIPTS[script]
    rng = np.random.default_rng(seed)
    nms = int(round(s['fs'] * 0.001))
    nx = (ms + 5) * nms
    x = rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)
    t = np.arange(nx)
    L = len(chips)
    mod = __import__('importlib').import_module
    code_rate = {1023: 1.023e6, 4092: 1.023e6, 2046: 2.046e6, 511: 0.511e6,
                 5115: 5.115e6}.get(L, None)
    if code_rate is None:
        code_rate = L * s['fs'] / s['n'] if True else 0
    incr = L / s['n'] if L != 10230 or s['n'] != 81920 else L / s['n']
    c = orc.resample_code(chips, 123.25, 0, incr, nx)
    fd = 137.0 + (s['carrier_step'] * key if s['fdma'] else 0.0)
    x += 3.0 * c * np.exp(2j * np.pi * fd * t / s['fs'])
    return x


@pytest.mark.parametrize('script,key,grid,ms', CASES, ids=[c[0] + '-%d' % i for i, c in enumerate(CASES)])
def test_oracle_matches_reference_search(script, key, grid, ms):
    ref_search, ns = ref_lift.lift_search(script)
    chips, mod = ref_chips(ns, script, key)
    x = synth(script, chips, key, ms, seed=zlib.crc32(script.encode()) % 1000)   # stable across processes (hash() is salted)
    got = orc.search_script(script, x, chips, key, grid, ms)
    want = ref_search(x, key, grid, ms)
    assert got == tuple(want), (got, want)      # same arithmetic -> bit-identical float64


def test_oracle_helpers_match_reference():
    rnco = ref_lift.ref_import('gnsstools.nco')
    rca = ref_lift.ref_import('gnsstools.gps.ca')
    f = -1234.5 / 4096000.0
    assert np.array_equal(orc.nco(f, 0, 8192), rnco.nco(f, 0, 8192))
    assert np.array_equal(orc.nco(f, 0.37, 1000), rnco.nco(f, 0.37, 1000))
    assert np.array_equal(orc.boc11(0, 0, 4092 / 32768., 32768), rnco.boc11(0, 0, 4092 / 32768., 32768))
    assert np.array_equal(orc.resample_code(rca.ca_code(5), 17.0, 0.3, 1023 / 4096., 4096),
                          rca.code(5, 17.0, 0.3, 1023 / 4096., 4096))
    rng = np.random.default_rng(3)
    for f, p in [(-9334875.0 / 69984000.0, 0), (0.01234, 0.25), (-0.49, 0.9)]:
        a = (rng.integers(-127, 128, 5000) + 1j * rng.integers(-127, 128, 5000)).astype(np.complex64)
        b = a.copy()
        orc.mix(a, f, p)
        rnco.mix(b, f, p)
        assert np.array_equal(a, b)


# --------------------------------------------------------------------------- serial long-code searches
def _serial_capture(chips, L, chip_rate, fs, n_total, phase_chips, fd, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 8, n_total) + 1j * rng.normal(0, 8, n_total)
    t = np.arange(n_total)
    c = orc.resample_code(chips, phase_chips, 0, chip_rate / fs, n_total)
    x += 2.0 * c * np.exp(2j * np.pi * fd * t / fs)
    return x.astype(np.complex64)


def test_oracle_matches_reference_l2cl_search():
    """acquire-gps-l2cl.py search() (75 hypotheses, hard-coded) at a small sample rate."""
    ref_search, ns = ref_lift.lift_search('gps-l2cl')
    l2cl = ns['l2cl']
    fs, ms, prn, doppler, cm_phase = 1.2e6, 40, 3, 431.0, 8317.2
    ns['fs'] = fs
    chips = np.asarray(l2cl.l2cl_code(prn))
    x = _serial_capture(chips, 767250, 511500.0, fs, int(fs * 0.001 * (ms + 5)), 17 * 10230 + cm_phase, doppler, seed=11)
    want = ref_search(x, prn, doppler, cm_phase, ms)
    got = orc.search_l2cl(x, chips, fs, doppler, cm_phase, ms)
    assert got == tuple(want) and got[1] == 17, (got, want)


@pytest.mark.parametrize('band,step', [('l1', 562500), ('l2', 437500)])
def test_oracle_matches_reference_glonass_p_search(band, step):
    """acquire-glonass-l{1,2}-p.py search() (1000 hypotheses, hard-coded) at a small sample rate."""
    ref_search, ns = ref_lift.lift_search('glonass-%s-p' % band)
    p = ns['p']
    fs, ms, chan, doppler, ca_phase = 5.0e6, 8, -2, 310.0, 278.6
    ns['fs'] = fs
    chips = np.asarray(p.p_code())
    x = _serial_capture(chips, 5110000, 5110000.0, fs, int(fs * 0.001 * (ms + 5)), 5110 * 421 + 10 * ca_phase,
                        step * chan + doppler, seed=12)
    want = ref_search(x, chan, doppler, ca_phase, ms)
    got = orc.search_glonass_p(x, chips, fs, step, chan, doppler, ca_phase, ms)
    assert got == tuple(want) and got[1] == 421, (got, want)


def test_serial_script_constants_match_reference_sources():
    """The constants acquire_serial.py hard-codes are the ones in the reference scripts."""
    import os
    src = {s: open(os.path.join(ref_lift.REF, 'acquire-%s.py' % s)).read() for s in ('gps-l2cl', 'glonass-l1-p', 'glonass-l2-p')}
    for token in ('range(75)', 'ms//20', 'int(fs*0.020)', '(k+block)*10230+l2cm_code_phase', "default=40"):
        assert token in src['gps-l2cl'], token
    for s, step in (('glonass-l1-p', '562500'), ('glonass-l2-p', '437500')):
        for token in ('range(1000)', 'ms//4', 'int(fs*0.004)', '5110*k + 10*ca_code_phase', 'cp += n*incr', '5110000.0/fs',
                      step + '*chan+doppler', "default=80"):
            assert token in src[s], (s, token)
