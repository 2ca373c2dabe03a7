"""Galileo E5a-I primary code (OS SIS ICD): two 14-stage registers, register 2 started per PRN,
truncated to 10230 chips. Surface of reference gnsstools/galileo/e5ai.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = 1.0 - 2.0 * _g.hex_to_bits('842E9', 20)      # CS20_1

e5ai_init = _g.icd_table('galileo.e5ai', 'e5ai_init')       # prn -> register-2 start state

_R1_TAPS = (13, 7, 5, 0)
_R2_TAPS = (13, 11, 7, 6, 4, 3)
r1 = _g.lfsr_fibonacci(14, _R1_TAPS, 0x3fff, code_length)

codes = {}


def make_e5ai(prn):
    return np.logical_xor(r1, _g.lfsr_fibonacci(14, _R2_TAPS, e5ai_init[prn], code_length))


def e5ai_code(prn):
    if prn not in codes:
        codes[prn] = make_e5ai(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e5ai_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
