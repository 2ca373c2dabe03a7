"""BeiDou B3I ranging code (BDS-SIS-ICD-B3I): two 13-stage registers; G1 is short-cycled
(reset to all ones after state 1111111111100), G2 starts from the PRN's initial state;
10230 chips. Surface of reference gnsstools/beidou/b3i.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = 1.0 - 2.0 * np.array([0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0])       # NH20

b3i_g2_initial = _g.icd_table('beidou.b3i', 'b3i_g2_initial')

_g1 = _g.stage(_g.lfsr_states(13, (0, 2, 3, 12), 0x1fff, code_length,
                              reset_from=_g.bits_to_int('1111111111100'), reset_to=0x1fff), 12)

codes = {}


def b3i(prn):
    g2 = _g.lfsr_fibonacci(13, (0, 4, 5, 6, 8, 9, 11, 12), b3i_g2_initial[prn], code_length)
    return np.logical_xor(_g1, g2).astype(np.float64)


def b3i_code(prn):
    if prn not in codes:
        codes[prn] = b3i(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b3i_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
