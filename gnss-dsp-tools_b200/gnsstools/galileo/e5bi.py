"""Galileo E5b-I primary code (OS SIS ICD): two 14-stage registers, register 2 started per PRN,
truncated to 10230 chips. Surface of reference gnsstools/galileo/e5bi.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = 1.0 - 2.0 * np.array([1, 1, 1, 0])          # CS4_1

e5bi_init = _g.icd_table('galileo.e5bi', 'e5bi_init')       # prn -> register-2 start state

_R1_TAPS = (13, 12, 10, 3)
_R2_TAPS = (13, 11, 8, 7, 4, 1)
r1 = _g.lfsr_fibonacci(14, _R1_TAPS, 0x3fff, code_length)

codes = {}


def make_e5bi(prn):
    return np.logical_xor(r1, _g.lfsr_fibonacci(14, _R2_TAPS, e5bi_init[prn], code_length))


def e5bi_code(prn):
    if prn not in codes:
        codes[prn] = make_e5bi(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e5bi_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
