"""Galileo E5a-Q primary code (OS SIS ICD): two 14-stage registers, register 2 started per PRN,
truncated to 10230 chips. Surface of reference gnsstools/galileo/e5aq.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = _g.secondary_table('galileo.e5aq')         # prn -> +-1, 100 chips (CS100)

e5aq_init = _g.icd_table('galileo.e5aq', 'e5aq_init')       # prn -> register-2 start state

_R1_TAPS = (13, 7, 5, 0)
_R2_TAPS = (13, 11, 7, 6, 4, 3)
r1 = _g.lfsr_fibonacci(14, _R1_TAPS, 0x3fff, code_length)

codes = {}


def make_e5aq(prn):
    return np.logical_xor(r1, _g.lfsr_fibonacci(14, _R2_TAPS, e5aq_init[prn], code_length))


def e5aq_code(prn):
    if prn not in codes:
        codes[prn] = make_e5aq(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e5aq_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
