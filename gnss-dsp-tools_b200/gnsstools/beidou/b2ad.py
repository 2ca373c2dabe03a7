"""BeiDou B2a data (B2ad) ranging code (BDS-SIS-ICD-B2a): two 13-stage registers, G1 from all ones and
re-initialised after chip 8189 (period 8190), G2 from the PRN's initial state; 10230 chips.
Surface of reference gnsstools/beidou/b2ad.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = 1.0 - 2.0 * np.array([0, 0, 0, 1, 0])

b2ad_g2_initial = _g.icd_table('beidou.b2ad', 'b2ad_g2_initial')     # prn -> 13-character bit string, stage 1 first

_G1_TAPS = (0, 4, 10, 12)
_G2_TAPS = (2, 4, 8, 10, 11, 12)
_g1 = _g.stage(_g.lfsr_states(13, _G1_TAPS, 0x1fff, code_length, reset_after=8189, reset_to=0x1fff), 12)

codes = {}


def b2ad(prn):
    g2 = _g.lfsr_fibonacci(13, _G2_TAPS, b2ad_g2_initial[prn], code_length)
    return np.logical_xor(_g1, g2).astype(np.float64)


def b2ad_code(prn):
    if prn not in codes:
        codes[prn] = b2ad(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b2ad_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
