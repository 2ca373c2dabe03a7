"""GLONASS standard-accuracy (C/A) code: 9-stage register 1 + x^5 + x^9, output of stage 7, 511 chips. FDMA — one code for every satellite, so code() has no PRN
argument. Surface of reference gnsstools/glonass/ca.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 511000
code_length = 511

c = _g.lfsr_fibonacci(9, (8, 4), (1 << 9) - 1, code_length, out_tap=6)


def ca_code():
    return c


def code(chips, frac, incr, n):
    return _g.resample(c, chips, frac, incr, n)


def correlate(x, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path)."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
