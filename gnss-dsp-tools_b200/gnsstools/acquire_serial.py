"""Serial long-code acquisitions on the GPU correlator bank — the B200 replacement for the
search() of acquire-gps-l2cl.py:18-33, acquire-glonass-l1-p.py:14-32 and acquire-glonass-l2-p.py:14-32.

The reference walks the hypotheses one by one: resample the long code at the hypothesised phase
(`code()`), multiply block, code and carrier, sum, accumulate |.| over the blocks, keep the first
maximum. Here every (hypothesis, block) sum is one CTA of `gnssacq_correlate_bank`; the host only
forms the float64 start phases exactly as the reference does (so every sample picks the same chip)
and scans the H metrics with the reference's strict '>'.
"""

import numpy as np

from . import _native


def _bank(x, chips01, nco_freq, n, blocks, base, incr, engine):
    eng = engine if engine is not None else _native.default_engine()
    if x is not None:
        x = np.asarray(x)
        if x.shape[0] < blocks * n:
            raise ValueError('capture too short: %d samples, search needs %d' % (x.shape[0], blocks * n))
        eng.set_signal(np.ascontiguousarray(x[:blocks * n], dtype=np.complex64))
    p = eng.correlate_bank(chips01, nco_freq, n, blocks, n, base, incr)
    return np.abs(p)


def _first_max(q):
    """m_metric,m_k = 0,0; `if q>m_metric` over ascending k (acquire-gps-l2cl.py:23-32)."""
    m_metric, m_k = 0, 0
    for k in range(q.shape[0]):
        v = 0
        for a in q[k]:                      # q = q + np.absolute(np.sum(p)), block by block
            v = v + a
        if v > m_metric:
            m_metric, m_k = v, k
    return m_metric, m_k


def search_l2cl(x, prn, doppler, l2cm_code_phase, ms, fs, engine=None, hypotheses=75):
    """search(x, prn, doppler, l2cm_code_phase, ms) of acquire-gps-l2cl.py (fs is the script's global).
    Returns (metric, k)."""
    from .gps import l2cl
    blocks = ms // 20
    n = int(fs * 0.020)
    if blocks <= 0:
        return 0, 0
    L = l2cl.code_length
    incr = l2cl.chip_rate / fs
    k = np.arange(hypotheses)[:, None]
    b = np.arange(blocks)[None, :]
    chips = (k + b) * 10230 + l2cm_code_phase                  # acquire-gps-l2cl.py:27
    base = (chips % L) + 0                                     # gnsstools/gps/l2cl.py:67: (chips%L) + frac
    q = _bank(x, l2cl.l2cl_code(prn), -doppler / fs, n, blocks, base, incr, engine)
    return _first_max(q)


def search_glonass_p(x, chan, doppler, ca_code_phase, ms, fs, carrier_step=562500, engine=None, hypotheses=1000):
    """search(x, chan, doppler, ca_code_phase, ms) of acquire-glonass-l1-p.py (carrier_step 562500)
    and acquire-glonass-l2-p.py (437500). Returns (metric, k)."""
    from .glonass import p
    blocks = ms // 4
    n = int(fs * 0.004)
    if blocks <= 0:
        return 0, 0
    incr = 5110000.0 / fs
    base = np.empty((hypotheses, blocks), dtype=np.float64)
    for k in range(hypotheses):
        cp = 5110 * k + 10 * ca_code_phase                     # acquire-glonass-l1-p.py:21
        for block in range(blocks):
            base[k, block] = (0 % p.code_length) + cp           # p.code(0, cp, ...): chips = 0, frac = cp
            cp += n * incr                                      # :27, float accumulation as the reference
    q = _bank(x, p.p_code(), -(carrier_step * chan + doppler) / fs, n, blocks, base, incr, engine)
    return _first_max(q)
