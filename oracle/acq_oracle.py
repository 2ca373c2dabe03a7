"""TEST INFRASTRUCTURE ONLY — numpy/scipy restatement of the reference's FFT
parallel-code-phase acquisition search (float64 end to end, like the reference).

Parity status: PINNED. The reference has no golden vectors for this path
(SURVEY.md §4, §8c), so the restatement is pinned against the reference's own
``search()`` functions executed in the build container (oracle/ref_lift.py lifts
them verbatim from /root/reference/acquire-*.py at fixture-generation time);
tests/test_oracle_vs_reference.py re-checks that whenever /root/reference is
present, and tests/golden/*.npz hold the reference outputs for the GPU box.

Each function cites the reference lines it follows. The third-party arithmetic
(scipy.fftpack fft/ifft, numpy absolute/argmax/mean) is unpinned upstream
(no requirements file); this container has scipy 1.18.1 / numpy 2.3.5.
"""

import numpy as np
import scipy.fftpack as fft

NT = 1024
# reference gnsstools/nco.py:3-4
nco_table = np.exp(2 * (np.pi) * (1j) * np.arange(NT) * (1.0 / NT))


def nco(f, p, n):
    """reference gnsstools/nco.py:6-10"""
    idx = p + f * np.arange(n)
    idx = np.floor(idx * NT).astype('int')
    idx = np.mod(idx, NT)
    return nco_table[idx]


def boc11(chips, frac, incr, n):
    """reference gnsstools/nco.py:12-19"""
    c = np.array([-1, 1])
    idx = (chips % 2) + frac + incr * np.arange(n)
    idx = np.floor(idx * 2).astype('int')
    return c[np.mod(idx, 2)]


def mix(x, f, p):
    """reference gnsstools/nco.py:30-41 (Numba loop) restated in closed form:
    dp_i = dp_0 + i*df in wrapping int64, index (dp_i >> 50) & 1023; the complex128
    product is rounded to x's dtype on store. Mutates x in place, returns None.
    The product is written out in real arithmetic (ac-bd, ad+bc, each operation rounded)
    because that is what Numba emits; numpy's own complex multiply uses fused
    multiply-adds on AVX-512 hosts and differs in the last bit where ac-bd cancels."""
    n = len(x)
    dp0 = np.int64(int(np.floor(p * NT * (1 << 50))))
    df = np.int64(int(np.floor(f * NT * (1 << 50))))
    with np.errstate(over='ignore'):
        dp = dp0 + df * np.arange(n, dtype=np.int64)      # wraps mod 2^64 like the loop
    idx = (dp >> np.int64(50)) & np.int64(NT - 1)
    t = nco_table[idx]
    a, b = x.real.astype(np.float64), x.imag.astype(np.float64)
    re = a * t.real - b * t.imag
    im = a * t.imag + b * t.real
    x.real = re
    x.imag = im


def resample_code(chip_bits, chips, frac, incr, n):
    """reference gnsstools/gps/ca.py:106-112 (same body in every generator module)"""
    L = len(chip_bits)
    idx = (chips % L) + frac + incr * np.arange(n)
    idx = np.floor(idx).astype('int')
    idx = np.mod(idx, L)
    return 1.0 - 2.0 * np.asarray(chip_bits, dtype=np.float64)[idx]


def doppler_bins(doppler_search):
    """reference acquire-gps-l1.py:26 — np.arange(min,max,incr), max excluded."""
    lo, hi, step = doppler_search
    return np.arange(lo, hi, step)


def replica(chip_bits, n, pad, boc, periods=1):
    """Time-domain replica before its FFT.
    circular: acquire-gps-l1.py:22-23; +BOC: acquire-gps-l1cd.py:23-26;
    zero-padded: acquire-gps-l5i.py:22-24; padded+BOC: acquire-galileo-e1b.py:23-26."""
    L = len(chip_bits)
    incr = float(periods * L) / n        # periods = 1 in every reference script
    c = resample_code(chip_bits, 0, 0, incr, n)
    if boc:
        c = c * boc11(0, 0, incr, n)
    if pad:
        c = np.concatenate((c, np.zeros(n)))
    return c


def search(x, chip_bits, fs, n, doppler_search, blocks, pad=False, boc=False,
           normalize=False, mod_L=False, carrier_hz=0.0, lag_limit=None,
           return_grid=False, periods=1):
    """One replica against the Doppler grid.

    Variant A (normalize=True):  acquire-gps-l1.py:18-40
    Variant B (carrier_hz!=0):   acquire-glonass-l1.py:18-39
    Variant C (boc, mod_L):      acquire-gps-l1cd.py:18-42
    Variant D (pad, mod_L):      acquire-gps-l5i.py:18-40, acquire-galileo-e1b.py:18-42

    `lag_limit` (not in the reference) restricts the argmax to the first
    `lag_limit` lags — used only for multi-period coherent configs whose exact
    alias ties make the fp64 argmax arbitrary; `periods` (not in the
    reference, always 1 there) is the number of code periods per coherent block.
    Returns (metric, code_chips, doppler_hz) and, with return_grid,
    also (idx, doppler_bin, q[D, N]).
    """
    L = len(chip_bits)
    N = 2 * n if pad else n
    c = fft.fft(replica(chip_bits, n, pad, boc, periods))
    m_metric, m_code, m_doppler = 0, 0, 0
    m_idx, m_bin = 0, -1
    grid = []
    for k, doppler in enumerate(doppler_bins(doppler_search)):
        q = np.zeros(N)
        w = nco(-(carrier_hz + doppler) / fs, 0, N)
        for block in range(blocks):
            b = x[(block * n):(block * n + N)]
            b = b * w
            r = fft.ifft(c * np.conj(fft.fft(b)))
            q = q + np.absolute(r)
        idx = np.argmax(q if lag_limit is None else q[:lag_limit])
        metric = q[idx] / np.mean(q) if normalize else q[idx]
        if metric > m_metric:
            m_metric = metric
            m_code = (periods * L) * (float(idx) / n)
            m_doppler = doppler
            m_idx, m_bin = int(idx), k
        if return_grid:
            grid.append(q)
    if mod_L:
        m_code = m_code % L
    if return_grid:
        return (m_metric, m_code, m_doppler), (m_idx, m_bin, np.array(grid))
    return m_metric, m_code, m_doppler


# ---------------------------------------------------------------------------
# Per-script constants (each row cites the script's search()).
# blocks(ms) is the non-coherent block count as a function of --time.
# ---------------------------------------------------------------------------

def _sig(module, fs, n, blocks, pad=False, boc=False, normalize=False, mod_L=False,
         carrier_step=0.0, fdma=False):
    return dict(module=module, fs=fs, n=n, blocks=blocks, pad=pad, boc=boc,
                normalize=normalize, mod_L=mod_L, carrier_step=carrier_step, fdma=fdma)


_ms = lambda ms: ms
SCRIPTS = {
    'gps-l1':       _sig('gps.ca', 4096000.0, 4096, _ms, normalize=True),
    'xona-x1':      _sig('xona.x1p', 4096000.0, 4096, _ms, normalize=True),
    'xona-x5p':     _sig('xona.x5p', 30690000.0, 30690, _ms, normalize=True),
    'glonass-l1':   _sig('glonass.ca', 16384000.0, 16384, _ms, carrier_step=562500.0, fdma=True),
    'glonass-l2':   _sig('glonass.ca', 16384000.0, 16384, _ms, carrier_step=437500.0, fdma=True),
    'gps-l1cd':     _sig('gps.l1cd', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True),
    'gps-l1cp':     _sig('gps.l1cp', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True),
    'beidou-b1cd':  _sig('beidou.b1cd', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True),
    'beidou-b1cp':  _sig('beidou.b1cp', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True),
    'galileo-e1b':  _sig('galileo.e1b', 8192000.0, 32768, lambda ms: ms // 4 - 1, pad=True, boc=True, mod_L=True),
    'galileo-e1c':  _sig('galileo.e1c', 8192000.0, 32768, lambda ms: ms // 4 - 1, pad=True, boc=True, mod_L=True),
    'beidou-b1i':   _sig('beidou.b1i', 8192000.0, 8192, _ms, pad=True, mod_L=True),
    'beidou-b2i':   _sig('beidou.b1i', 8192000.0, 8192, _ms, pad=True, mod_L=True),
    'gps-l2cm':     _sig('gps.l2cm', 4096000.0, 81920, lambda ms: ms // 20 - 1, pad=True, mod_L=True),
    'beidou-b2ad':  _sig('beidou.b2ad', 30690000.0, 30690, lambda ms: 80, pad=True, mod_L=True),
    'galileo-e6b':  _sig('galileo.e6b', 15345000.0, 15345, _ms, pad=True, mod_L=True),
    'galileo-e6c':  _sig('galileo.e6c', 15345000.0, 15345, _ms, pad=True, mod_L=True),
}
for _name, _mod in [('gps-l5i', 'gps.l5i'), ('gps-l5q', 'gps.l5q'),
                    ('galileo-e5ai', 'galileo.e5ai'), ('galileo-e5aq', 'galileo.e5aq'),
                    ('galileo-e5bi', 'galileo.e5bi'), ('galileo-e5bq', 'galileo.e5bq'),
                    ('beidou-b2ap', 'beidou.b2ap'), ('beidou-b2bi', 'beidou.b2bi'),
                    ('beidou-b2bq', 'beidou.b2bq'), ('beidou-b3i', 'beidou.b3i'),
                    ('glonass-l3ocd', 'glonass.l3ocd'), ('glonass-l3ocp', 'glonass.l3ocp')]:
    SCRIPTS[_name] = _sig(_mod, 30690000.0, 30690, _ms, pad=True, mod_L=True)


def search_script(script, x, chip_bits, key, doppler_search, ms, **kw):
    """search() of reference acquire-<script>.py; `key` is the PRN, or the FDMA
    channel for the GLONASS scripts (acquire-glonass-l1.py:28)."""
    s = SCRIPTS[script]
    carrier = s['carrier_step'] * key if s['fdma'] else 0.0
    return search(x, chip_bits, s['fs'], s['n'], doppler_search, s['blocks'](ms),
                  pad=s['pad'], boc=s['boc'], normalize=s['normalize'],
                  mod_L=s['mod_L'], carrier_hz=carrier, **kw)


# --------------------------------------------------------------------------- serial long-code searches
# (the time-domain correlator bank behind gnssacq_correlate_bank)

def search_l2cl(x, chip_bits, fs, doppler, l2cm_code_phase, ms, hypotheses=75):
    """reference acquire-gps-l2cl.py:18-33: 75 hypotheses of which 10230-chip L2CM period the
    767250-chip L2CL code is in, 20 ms blocks, q = sum_block |sum(x*c*w)|, first maximum by strict '>'."""
    blocks = ms // 20
    n = int(fs * 0.020)
    w = nco(-doppler / fs, 0, n)
    incr = 511500 / fs                                   # l2cl.chip_rate
    m_metric, m_k = 0, 0
    for k in range(hypotheses):
        q = 0
        for block in range(blocks):
            c = resample_code(chip_bits, (k + block) * 10230 + l2cm_code_phase, 0, incr, n)
            p = x[n * block:n * (block + 1)] * c * w
            q = q + np.absolute(np.sum(p))
        if q > m_metric:
            m_metric = q
            m_k = k
    return m_metric, m_k


def search_glonass_p(x, chip_bits, fs, carrier_step, chan, doppler, ca_code_phase, ms, hypotheses=1000):
    """reference acquire-glonass-l1-p.py:14-32 (carrier_step 562500) and acquire-glonass-l2-p.py:14-32
    (437500): 1000 hypotheses of which C/A period (5110 P chips) the 5110000-chip P code is in, 4 ms
    blocks; the code phase advances by the float accumulation cp += n*incr and enters code() as `frac`."""
    blocks = ms // 4
    n = int(fs * 0.004)
    w = nco(-(carrier_step * chan + doppler) / fs, 0, n)
    m_metric, m_k = 0, 0
    for k in range(hypotheses):
        q = 0
        cp = 5110 * k + 10 * ca_code_phase
        for block in range(blocks):
            incr = 5110000.0 / fs
            c = resample_code(chip_bits, 0, cp, incr, n)
            xp = x[n * block:n * (block + 1)] * c * w
            q = q + np.absolute(np.sum(xp))
            cp += n * incr
        if q > m_metric:
            m_metric = q
            m_k = k
    return m_metric, m_k
