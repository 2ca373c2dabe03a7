"""Galileo E5b-Q primary code (OS SIS ICD): two 14-stage registers, register 2 started per PRN,
truncated to 10230 chips. Surface of reference gnsstools/galileo/e5bq.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

secondary_code = _g.secondary_table('galileo.e5bq')         # prn -> +-1, 100 chips (CS100)

e5bq_init = _g.icd_table('galileo.e5bq', 'e5bq_init')       # prn -> register-2 start state

_R1_TAPS = (13, 12, 10, 3)
_R2_TAPS = (13, 9, 8, 5, 4, 0)
r1 = _g.lfsr_fibonacci(14, _R1_TAPS, 0x3fff, code_length)

codes = {}


def make_e5bq(prn):
    return np.logical_xor(r1, _g.lfsr_fibonacci(14, _R2_TAPS, e5bq_init[prn], code_length))


def e5bq_code(prn):
    if prn not in codes:
        codes[prn] = make_e5bq(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e5bq_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
