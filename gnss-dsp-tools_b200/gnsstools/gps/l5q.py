"""GPS L5 Q5 code (IS-GPS-705): XA (13 stages, short-cycled to 8190) xor XB advanced by the
PRN's offset; 10230 chips. Surface of reference gnsstools/gps/l5q.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

nh_code = [0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0]                                              # NH20

l5q_init = _g.icd_table('gps.l5q', 'l5q_init')         # prn -> XB advance

# XA: taps 9,10,12,13, reset to all ones after state 1111111111101; XB: taps 1,3,4,6,7,8,12,13
xa = _g.stage(_g.lfsr_states(13, (12, 11, 9, 8), 0x1fff, code_length, reset_from=0x17ff, reset_to=0x1fff), 12)
xb = _g.lfsr_fibonacci(13, (12, 11, 7, 6, 5, 3, 2, 0), 0x1fff, 8191)

codes = {}


def make_l5q(prn):
    return np.logical_xor(xa, xb[(l5q_init[prn] + np.arange(code_length)) % 8191])


def l5q_code(prn):
    if prn not in codes:
        codes[prn] = make_l5q(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(l5q_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
