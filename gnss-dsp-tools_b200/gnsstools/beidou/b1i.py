"""BeiDou B1I / B2I ranging code (BDS-SIS-ICD-B1I): two 11-stage registers from 01010101010,
G2 output taken as the xor of the PRN's phase-selection taps; 2046 chips.
Surface of reference gnsstools/beidou/b1i.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 2046000
code_length = 2046

secondary_code = 1.0 - 2.0 * np.array([0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0])       # NH20

b1i_g2_taps = _g.icd_table('beidou.b1i', 'b1i_g2_taps')     # prn -> 1-based G2 stages to xor

_START = _g.bits_to_int([0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0])
_g1 = _g.lfsr_fibonacci(11, (0, 6, 7, 8, 9, 10), _START, code_length)
_g2_states = _g.lfsr_states(11, (0, 1, 2, 3, 4, 7, 8, 10), _START, code_length)

codes = {}


def b1i(prn):
    c = _g1.astype(np.int64)
    for tap in b1i_g2_taps[prn]:
        c ^= (_g2_states >> (tap - 1)) & 1
    return c.astype(np.float64)


def b1i_code(prn):
    if prn not in codes:
        codes[prn] = b1i(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b1i_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
