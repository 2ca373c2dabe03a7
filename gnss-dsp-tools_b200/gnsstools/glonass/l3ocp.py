"""GLONASS L3OC pilot (L3OCp) code (CDMA ICD): 14-stage register from 00110100111000 xor a 7-stage
register loaded with the PRN number + 64, MSB first; 10230 chips. PRN range 0-63.
Surface of reference gnsstools/glonass/l3ocp.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = 1.0 - 2.0 * np.array([0, 0, 0, 0, 1, 1, 0, 1, 0, 1])

_g2 = _g.lfsr_fibonacci(14, (13, 12, 7, 3), _g.bits_to_int('00110100111000'), code_length)

codes = {}


def make_l3ocp(n):
    start = [((n + 64) >> (6 - i)) & 1 for i in range(7)]
    return np.logical_xor(_g.lfsr_fibonacci(7, (6, 5), start, code_length), _g2).astype(np.float64)


def l3ocp_code(n):
    if n not in codes:
        codes[n] = make_l3ocp(n)
    return codes[n]


def code(prn, chips, frac, incr, n):
    return _g.resample(l3ocp_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
